// ResNet101 extractor (resnet.cu).  Internal header.
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "common.h"

namespace vg {

struct ResNet {
  // torchvision resnet101 (Bottleneck v1.5, blocks 3/4/23/3, widths 64/128/256/512, expansion 4) as built by the reference's
  // Backbone (vgqa/core/vision/backbone.py:104-113): FrozenBatchNorm2d (:13-57, eps 1e-5), no dilation, output = layer4
  static constexpr int kLayers = 4;
  int blocks[kLayers] = {3, 4, 23, 3};
  struct Conv {            // a convolution with its FrozenBN folded in: W [O, taps * C] bf16 (k = tap * C + c), bias [O] fp32
    bf16* W = nullptr;
    float* bias = nullptr;
    int O = 0, C = 0, taps = 1;
  };
  struct Block { Conv c1, c2, c3, down; bool has_down = false; int stride = 1; };
  Conv stem;               // 7x7 / 2: W [64, 192] (k = c * 49 + ky * 7 + kx, columns 147..191 zero)
  std::vector<Block> layer[kLayers];
  bool loaded = false;

  // workspace for one chunk of frames (allocated on first use, grown on demand)
  int cap_frames = 0, cap_R = 0;
  bf16 *a0 = nullptr, *s0 = nullptr, *x = nullptr, *y = nullptr, *idn = nullptr, *o1 = nullptr, *o2 = nullptr, *a2 = nullptr, *xs = nullptr;
  int launches = 0;

  using GetFn = std::function<const float*(const std::string&, std::vector<int64_t>)>;
  using HasFn = std::function<bool(const std::string&)>;
  void pack(const HasFn& has, const GetFn& get, const std::function<bf16*(const float*, size_t)>& to_bf16,
            const std::function<float*(const float*, size_t)>& to_f32);
  void ensure_workspace(int frames, int R);
  void release();
  // frames NCHW fp32 [n, 3, R, R] → layer4 map channels-last [n, R/32, R/32, 2048] (bf16 and / or fp32); layer_out[l] (optional):
  // the output of layer l + 1, channels-last fp32 [n, R / (4 << l), R / (4 << l), 256 << l]
  int forward(const float* frames, int n, int R, bf16* out_bf16, float* out_f32, float* const* layer_out, cudaStream_t st);
};

}  // namespace vg
