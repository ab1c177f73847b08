// Video-Swin-T extractor (swin.cu).  Internal header.
#pragma once
#include <functional>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "common.h"

namespace vg {

struct SwinNet {
  // configs['video_swin_t_p4w7'] (video_swin_transformer.py:688-700): patch (1,4,4), embed 96, depths 2/2/6/2, heads 3/6/12/24,
  // window (8,7,7), mlp ratio 4
  static constexpr int kStages = 4;
  int embed = 96, wd = 8, wh = 7, ww = 7;
  int depths[kStages] = {2, 2, 6, 2};
  int heads[kStages] = {3, 6, 12, 24};
  struct Block {
    float *n1w, *n1b, *n2w, *n2b, *bqkv, *bproj, *bfc1, *bfc2;
    bf16 *Wqkv, *Wproj, *Wfc1, *Wfc2;                                    // channel-padded to multiples of 64
    bf16* sbias;                                                         // [heads][N][N] = relative position bias / scale
  };
  struct Stage {
    int C = 0, Cp = 0, Nqkv = 0, heads = 0;   // channels, padded channel stride, padded packed q|k|v width
    std::vector<Block> blocks;
    bool loaded = false;
    // PatchMerging INTO the next stage (downsamples[s]): LayerNorm(4C) + Linear(4C → 2C, no bias)
    float *mnw = nullptr, *mnb = nullptr;
    bf16* Wred = nullptr;
  } st[kStages];
  // PatchEmbed3D: Conv3d(3, 96, (1,4,4)) as a [128 x 64] GEMM + LayerNorm(96)
  bf16* Wpe = nullptr;
  float *bpe = nullptr, *pnw = nullptr, *pnb = nullptr;
  bool full = false;   // patch embedding + all stages + PatchMerging weights present (else: last stage only)

  // region-id tables of the shifted windows per (D, H, W) of a stage: rid [nsets][N] + gset [windows per clip], device-resident
  struct MaskTab { uint8_t* rid = nullptr; uint8_t* gset = nullptr; int groups_cap = 0; int nW = 0; std::vector<uint8_t> h_rid, h_gset; };
  std::map<std::tuple<int, int, int, int>, MaskTab> masks;   // key: D, H, W, clips

  // workspace (allocated on first use, grown on demand; not part of the forward's arena)
  float *x32 = nullptr, *y32 = nullptr;
  bf16 *xn = nullptr, *xw = nullptr, *qkv = nullptr, *ao = nullptr, *hid = nullptr, *xm = nullptr, *a0 = nullptr;
  size_t cap_units = 0, cap_qkv = 0, cap_frames = 0;   // capacities: (rows x padded channels) of the largest stage, packed q|k|v elements, patch rows
  int launches = 0;

  using GetFn = std::function<const float*(const std::string&, std::vector<int64_t>)>;
  using HasFn = std::function<bool(const std::string&)>;
  void pack(const HasFn& has, const GetFn& get, const std::function<bf16*(const float*, size_t)>& to_bf16,
            const std::function<float*(const float*, size_t)>& to_f32);
  void ensure_workspace(size_t units, size_t qkv_units, size_t frame_rows);
  void stage_units(int s, int clips, int D, int H, int W, size_t& units, size_t& qkv_units) const;
  void release();
  // one stage on the channels-last fp32 stream x32 [clips, D, H, W, Cp] (in place)
  void run_stage(int s, int clips, int D, int H, int W, cudaStream_t st);
  // last stage only: x channels-last fp32 [clips, D, H, W, 768]
  int forward_stage4(const float* x, int clips, int D, int H, int W, bf16* out_bf16, float* out_f32, cudaStream_t st);
  // whole extractor: frames NCHW fp32 [clips*T, 3, R, R]; stage_out[s] (optional) receive the stage outputs channels-last fp32
  int forward_full(const float* frames, int clips, int T, int R, bf16* out_bf16, float* out_f32, float* const* stage_out, cudaStream_t st);
};

}  // namespace vg
