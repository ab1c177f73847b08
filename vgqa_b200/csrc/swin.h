// Last stage of the Video-Swin-T extractor (swin.cu).  Internal header.
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "common.h"

namespace vg {

struct SwinStage {
  int dim = 768, heads = 24, wd = 8, wh = 7, ww = 7, depth = 2;   // configs['video_swin_t_p4w7'] (video_swin_transformer.py:688-700)
  struct Block {
    float *n1w, *n1b, *n2w, *n2b, *bqkv, *bproj, *bfc1, *bfc2, *sbias;
    bf16 *Wqkv, *Wproj, *Wfc1, *Wfc2;
    int bias_sets;
  };
  std::vector<Block> blocks;
  bool loaded = false;
  // workspace (allocated on first use, grown on demand; not part of the forward's arena)
  float *x32 = nullptr, *y32 = nullptr;
  bf16 *xn = nullptr, *xs = nullptr, *qkv = nullptr, *ao = nullptr, *hid = nullptr;
  size_t cap_rows = 0;
  int launches = 0;

  void pack(const std::function<const float*(const std::string&, std::vector<int64_t>)>& get,
            const std::function<bf16*(const float*, size_t)>& to_bf16, const std::function<float*(const float*, size_t)>& to_f32);
  void ensure_workspace(size_t rows);
  void release();
  int forward(const float* x, int clips, int D, int H, int W, bf16* out_bf16, float* out_f32, cudaStream_t st);
};

}  // namespace vg
