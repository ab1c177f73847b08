// The last stage of the Video-Swin-T extractor on this library's kernels (first piece):
// `BasicLayer` = 2 x SwinTransformerBlock3D, dim 768, 24 heads of 32, window (8,7,7), mlp ratio 4 — the module behind
// `vid.layers[3]` of VSTGNet (vgqa/core/vision/video_swin_transformer.py:176-275,337-398; grounding_net.py:67-71,104-105).
//
//   per block:   x += proj( W-MSA( LN1(x) ) )          window attention with the relative position bias (:143-165); odd blocks
//                                                      on the map rolled by 4 frames, with the -100 mask between the two
//                                                      halves of the wrapped window (compute_mask, :311-325)
//                x += fc2( gelu( fc1( LN2(x) ) ) )     erf GELU
//
// A 224 px clip leaves a 7x7 map at this stage, so a window is 8 consecutive frames x all 49 positions = 392 CONTIGUOUS token
// rows of the channels-last map: the window partition is free, the temporal roll is a row gather.  Kernels: LayerNorm rows
// (text_tower.cu), the tcgen05 GEMMs (qkv 768 → 2304, proj + fp32 residual, fc1 + GELU, fc2 + fp32 residual) and the multi-tile
// tcgen05 attention of attn_tc_long.cu generalised to 24 heads and an additive score term (bias / scale, fp32, L2-resident).
// The fp32 residual stream stays fp32 as in the encoder.  Output: channels-last bf16 = exactly the `vid_raw` / raw_layout = 1
// input of vgqa_forward (input_proj2 reads it as its GEMM operand), and optionally fp32.
#include <cmath>
#include <string>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"
#include "swin.h"

namespace vg {

// dst[(b, d, :)] = src[(b, (d + shift) mod D, :)]  — torch.roll(x, -shift, dims=1) on frames of `frame_elems` bf16 (16-byte chunks)
__global__ void __launch_bounds__(256) roll_frames_bf16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int D,
                                                               long long chunks_per_frame, int shift, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long fr = i / chunks_per_frame, off = i - fr * chunks_per_frame;
    const long long b = fr / D;
    const int d = (int)(fr - b * D);
    dst[i] = __ldg(src + (b * D + (d + shift) % D) * chunks_per_frame + off);
  }
}
// x[(b, d, :)] += y[(b, (d - shift) mod D, :)]  — the reverse roll fused with the residual add (fp32, 16-byte chunks)
__global__ void __launch_bounds__(256) add_rolled_f32_kernel(float4* __restrict__ x, const float4* __restrict__ y, int D,
                                                             long long chunks_per_frame, int shift, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long fr = i / chunks_per_frame, off = i - fr * chunks_per_frame;
    const long long b = fr / D;
    const int d = (int)(fr - b * D);
    const float4 a = x[i], q = __ldg(y + (b * D + (d - shift + D) % D) * chunks_per_frame + off);
    x[i] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
  }
}
__global__ void __launch_bounds__(256) f32_to_bf16_rows_kernel(const float4* __restrict__ x, uint2* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(x + i);
    y[i] = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
  }
}

static int grid_for(long long n) { return (int)std::min<long long>((n + 255) / 256, 148 * 16); }

void SwinStage::pack(const std::function<const float*(const std::string&, std::vector<int64_t>)>& get,
                     const std::function<bf16*(const float*, size_t)>& to_bf16, const std::function<float*(const float*, size_t)>& to_f32) {
  const int nb = (2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1), N = wd * wh * ww;
  const float inv_scale = std::sqrt((float)(dim / heads));
  // relative position index (video_swin_transformer.py:97-112) and the shift-mask groups of the wrapped window (:311-325)
  std::vector<int> idx((size_t)N * N);
  for (int a = 0; a < N; ++a)
    for (int b = 0; b < N; ++b) {
      const int da = a / (wh * ww), ha = (a / ww) % wh, wa = a % ww, db = b / (wh * ww), hb = (b / ww) % wh, wb = b % ww;
      idx[(size_t)a * N + b] = (da - db + wd - 1) * (2 * wh - 1) * (2 * ww - 1) + (ha - hb + wh - 1) * (2 * ww - 1) + (wa - wb + ww - 1);
    }
  blocks.resize(depth);
  for (int i = 0; i < depth; ++i) {
    const std::string p = "vid.layers.3.blocks." + std::to_string(i) + ".";
    Block& k = blocks[i];
    auto lin = [&](const std::string& n, int o, int in, bf16*& W, float*& b) {
      W = to_bf16(get(p + n + ".weight", {o, in}), (size_t)o * in);
      b = to_f32(get(p + n + ".bias", {o}), o);
    };
    k.n1w = to_f32(get(p + "norm1.weight", {dim}), dim); k.n1b = to_f32(get(p + "norm1.bias", {dim}), dim);
    k.n2w = to_f32(get(p + "norm2.weight", {dim}), dim); k.n2b = to_f32(get(p + "norm2.bias", {dim}), dim);
    lin("attn.qkv", 3 * dim, dim, k.Wqkv, k.bqkv);
    lin("attn.proj", dim, dim, k.Wproj, k.bproj);
    lin("mlp.fc1", 4 * dim, dim, k.Wfc1, k.bfc1);
    lin("mlp.fc2", dim, 4 * dim, k.Wfc2, k.bfc2);
    const float* tab = get(p + "attn.relative_position_bias_table", {nb, heads});
    const int sets = (i % 2 == 1) ? 2 : 1;
    std::vector<float> sb((size_t)sets * heads * N * N);
    for (int s = 0; s < sets; ++s)
      for (int h = 0; h < heads; ++h)
        for (int a = 0; a < N; ++a)
          for (int b = 0; b < N; ++b) {
            float v = tab[(size_t)idx[(size_t)a * N + b] * heads + h];
            if (s == 1) {   // wrapped window of the rolled map: frames [0, wd - shift) and [wd - shift, wd) belong to different groups
              const int ga = (a / (wh * ww)) < wd - wd / 2, gb = (b / (wh * ww)) < wd - wd / 2;
              if (ga != gb) v += -100.0f;
            }
            sb[(((size_t)s * heads + h) * N + a) * N + b] = v * inv_scale;
          }
    k.sbias = to_f32(sb.data(), sb.size());
    k.bias_sets = sets;
  }
  loaded = true;
}

void SwinStage::ensure_workspace(size_t rows) {
  if (rows <= cap_rows) return;
  for (void* q : {(void*)x32, (void*)y32, (void*)xn, (void*)xs, (void*)qkv, (void*)ao, (void*)hid})
    if (q) cudaFree(q);
  const size_t C = dim;
  VG_CUDA(cudaMalloc(&x32, rows * C * 4)); VG_CUDA(cudaMalloc(&y32, rows * C * 4));
  VG_CUDA(cudaMalloc(&xn, rows * C * 2)); VG_CUDA(cudaMalloc(&xs, rows * C * 2));
  VG_CUDA(cudaMalloc(&qkv, rows * 3 * C * 2)); VG_CUDA(cudaMalloc(&ao, rows * C * 2)); VG_CUDA(cudaMalloc(&hid, rows * 4 * C * 2));
  cap_rows = rows;
}

void SwinStage::release() {
  for (void* q : {(void*)x32, (void*)y32, (void*)xn, (void*)xs, (void*)qkv, (void*)ao, (void*)hid})
    if (q) cudaFree(q);
  x32 = y32 = nullptr; xn = xs = qkv = ao = hid = nullptr; cap_rows = 0;
}

// x: channels-last fp32 [clips, D, H, W, dim]; out_bf16 (channels-last bf16) and / or out_f32 receive the stage output
int SwinStage::forward(const float* x, int clips, int D, int H, int W, bf16* out_bf16, float* out_f32, cudaStream_t st) {
  VG_CHECK(loaded, "vgqa_swin_stage needs the 'vid.layers.3.blocks.*' weights");
  VG_CHECK(clips >= 1 && H == wh && W == ww, "vgqa_swin_stage: the map must fill the window in H and W (7x7: 224 px clips)");
  const int w_d = D < wd ? D : wd;
  VG_CHECK(D >= 1 && D % w_d == 0, "vgqa_swin_stage: the number of frames must be a multiple of the temporal window (8)");
  VG_CHECK(w_d * H * W > 128, "vgqa_swin_stage: windows of at most 128 tokens are not supported (needs at least 3 frames)");
  VG_CHECK(w_d == wd, "vgqa_swin_stage: clips shorter than the temporal window (8 frames) are not supported");
  const int shift = D > wd ? wd / 2 : 0;                    // get_window_size: no shift along a clamped axis (:53-66)
  const size_t P = (size_t)H * W, rows = (size_t)clips * D * P, C = dim;
  const int N = w_d * H * W, groups = clips * (D / w_d), wpc = D / w_d;
  const int launches0 = launches;
  ensure_workspace(rows);
  VG_CUDA(cudaMemcpyAsync(x32, x, rows * C * 4, cudaMemcpyDeviceToDevice, st));
  const long long cpf_bf16 = (long long)P * C / 8, cpf_f32 = (long long)P * C / 4;
  for (int i = 0; i < depth; ++i) {
    Block& k = blocks[i];
    const int sh = (i % 2 == 1) ? shift : 0;
    ln_rows_wide(x32, k.n1w, k.n1b, 1e-5f, nullptr, xn, (int)rows, (int)C, st);
    const bf16* a = xn;
    if (sh) {
      const long long n = (long long)rows * C / 8;
      roll_frames_bf16_kernel<<<grid_for(n), 256, 0, st>>>(reinterpret_cast<const uint4*>(xn), reinterpret_cast<uint4*>(xs), D, cpf_bf16, sh, n);
      a = xs; ++launches;
    }
    { GemmEpi ep; ep.C = qkv; ep.ldc = 3 * (int)C; ep.bias = k.bqkv; ep.bias_ld = 3 * (int)C;
      gemm_bf16_tn(a, (int)C, k.Wqkv, (int)C, (int)rows, 3 * (int)C, (int)C, ep, st); }
    window_attn_tc(qkv, ao, groups, N, heads, k.sbias, sh ? k.bias_sets : 1, wpc, 1.0f / std::sqrt((float)(C / heads)), st);
    if (sh) {   // proj → fp32, then the reverse roll fused with the residual add
      GemmEpi ep; ep.C = y32; ep.ldc = (int)C; ep.c_f32 = 1; ep.bias = k.bproj; ep.bias_ld = (int)C;
      gemm_bf16_tn(ao, (int)C, k.Wproj, (int)C, (int)rows, (int)C, (int)C, ep, st);
      const long long n = (long long)rows * C / 4;
      add_rolled_f32_kernel<<<grid_for(n), 256, 0, st>>>(reinterpret_cast<float4*>(x32), reinterpret_cast<const float4*>(y32), D, cpf_f32, sh, n);
      ++launches;
    } else {
      GemmEpi ep; ep.C = y32; ep.ldc = (int)C; ep.c_f32 = 1; ep.bias = k.bproj; ep.bias_ld = (int)C; ep.res32 = x32; ep.ldres32 = (int)C;
      gemm_bf16_tn(ao, (int)C, k.Wproj, (int)C, (int)rows, (int)C, (int)C, ep, st);
      std::swap(x32, y32);
    }
    ln_rows_wide(x32, k.n2w, k.n2b, 1e-5f, nullptr, xn, (int)rows, (int)C, st);
    { GemmEpi ep; ep.C = hid; ep.ldc = 4 * (int)C; ep.bias = k.bfc1; ep.bias_ld = 4 * (int)C; ep.act = ACT_GELU;
      gemm_bf16_tn(xn, (int)C, k.Wfc1, (int)C, (int)rows, 4 * (int)C, (int)C, ep, st); }
    { GemmEpi ep; ep.C = y32; ep.ldc = (int)C; ep.c_f32 = 1; ep.bias = k.bfc2; ep.bias_ld = (int)C; ep.res32 = x32; ep.ldres32 = (int)C;
      gemm_bf16_tn(hid, 4 * (int)C, k.Wfc2, 4 * (int)C, (int)rows, (int)C, 4 * (int)C, ep, st); }
    std::swap(x32, y32);
    launches += 7;
  }
  VG_CUDA(cudaGetLastError());
  if (out_f32) VG_CUDA(cudaMemcpyAsync(out_f32, x32, rows * C * 4, cudaMemcpyDeviceToDevice, st));
  if (out_bf16) {
    const long long n = (long long)rows * C / 4;
    f32_to_bf16_rows_kernel<<<grid_for(n), 256, 0, st>>>(reinterpret_cast<const float4*>(x32), reinterpret_cast<uint2*>(out_bf16), n);
    ++launches;
  }
  VG_CUDA(cudaGetLastError());
  return launches - launches0;
}

}  // namespace vg
