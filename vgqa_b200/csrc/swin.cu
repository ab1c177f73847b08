// The Video-Swin-T extractor on this library's kernels: `self.vid` of VSTGNet = VideoSwinTransformerBackbone
// (vgqa/core/vision/video_swin_transformer.py:626-685; grounding_net.py:67-71,104-105) — PatchEmbed3D with patch (1,4,4) +
// LayerNorm (:403-443), four stages of SwinTransformerBlock3D (depths 2/2/6/2, dims 96/192/384/768, heads 3/6/12/24 of 32, window
// (8,7,7), mlp ratio 4; :176-275,337-398) with PatchMerging between them (:278-308).  Either the whole extractor (frames in, the
// last stage's map out) or the last stage alone (`vid.layers[3]`), depending on which weights were given.
//
//   per block:   x += proj( W-MSA( LN1(x) ) )          window attention with the relative position bias (:143-165); odd blocks on
//                                                      the cyclically rolled map with the -100 mask of compute_mask (:311-325)
//                x += fc2( gelu( fc1( LN2(x) ) ) )     erf GELU
//
// Layout: channels-last fp32 residual stream x32 [clips, D, H, W, Cp] (Cp = channels padded to a multiple of 64: 96 → 128 in the
// first stage, zero in the pad columns), exactly as in the encoder.  Kernels:
//   * LayerNorm rows (one warp per token) → bf16 GEMM operand;
//   * window partition + cyclic roll ride on LayerNorm 1 (one pass: fp32 map row in, normalised bf16 row out at its window position);
//     window reverse + un-roll + residual add ride on LayerNorm 2 (x[m] += y[r], LN of the sum in the same pass); where a window is
//     a run of whole frames and nothing is rolled (last stage at 7x7, even blocks) the partition is free;
//   * the tcgen05 GEMMs of the encoder (qkv, proj → fp32, fc1 + GELU, fc2 + fp32 residual, patch embedding, PatchMerging reduction);
//   * the multi-tile tcgen05 attention of attn_tc_long.cu with `heads` heads, the bias table [heads][392][392] (divided by the
//     scale, L2-resident) and the shift mask from per-token region ids (a few hundred bytes per distinct window kind);
//   * PatchMerging: 2x2 gather + LayerNorm(4C) in one kernel, then the reduction GEMM.
// Map sides that are not multiples of the window take the reference's padding path (zero tokens at the end of every axis, for the
// attention half of a block only); supported: T >= 8 and frame sides that are multiples of 32 from 224 on (no clamped windows).
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"
#include "swin.h"

namespace vg {

// ------------------------------------------------------------------------------------------------ kernels
// D, H, W: the map; Dp, Hp, Wp: its sides rounded up to multiples of the window (the reference zero-pads LN1(x) at the END of every
// axis, :205-211); the windows and the cyclic roll live on the padded grid
struct WinGeom { int B, D, H, W, Dp, Hp, Wp, wd, wh, ww, sd, sh, sw; };
// window-order row (padded grid) → row of the map, or -1 for a padding token; 32-bit arithmetic: rows < 2^31
__device__ __forceinline__ int win_src_row(const WinGeom& g, int r) {
  const unsigned N = g.wd * g.wh * g.ww;
  const unsigned grp = (unsigned)r / N, n = (unsigned)r - grp * N;
  const unsigned nWh = g.Hp / g.wh, nWw = g.Wp / g.ww, nWd = g.Dp / g.wd;
  const unsigned q1 = grp / nWw, wwi = grp - q1 * nWw, q2 = q1 / nWh, hwi = q1 - q2 * nWh, b = q2 / nWd, dwi = q2 - b * nWd;
  const unsigned t1 = n / g.ww, wl = n - t1 * g.ww, dd = t1 / g.wh, hh = t1 - dd * g.wh;
  unsigned d = dwi * g.wd + dd + g.sd, h = hwi * g.wh + hh + g.sh, w = wwi * g.ww + wl + g.sw;   // the shifts are smaller than the sides
  if (d >= (unsigned)g.Dp) d -= g.Dp;
  if (h >= (unsigned)g.Hp) h -= g.Hp;
  if (w >= (unsigned)g.Wp) w -= g.Wp;
  if (d >= (unsigned)g.D || h >= (unsigned)g.H || w >= (unsigned)g.W) return -1;
  return (int)(((b * g.D + d) * g.H + h) * g.W + w);
}
// LayerNorm fused with the window plumbing (rows r in WINDOW order; m = the map row that r comes from / goes to):
//   MODE 1  (LN1 + window_partition(roll(pad(x), -shift))):   y16[r] = LN(x[m]), zeros for a padding token
//   MODE 2  (x += crop(roll(window_reverse(y), +shift)), then LN2):   x[m] += add[r];  y16[m] = LN(x[m]); padding tokens are skipped
//   MODE 0 / 3  (no window map, m = r):   y16[r] = LN(x[r])  /  x[r] = LN(x[r]) in place, fp32 (the patch embedding's norm)
// Saves the bf16 round trip of a separate gather (MODE 1) and LayerNorm's re-read of the fp32 row (MODE 2).  A warp walks rows with a
// grid stride and keeps RW rows (NJ float4 per lane each) in flight: one 512-byte row per warp at a time leaves the loads latency-bound
// (2.3 TB/s measured on the 96-channel stage).
template <int MODE, int NJ, int RW>
__global__ void __launch_bounds__(256, 4) ln_window_kernel(float* x, int ld, int C, const float* __restrict__ w, const float* __restrict__ b,
                                                        float eps, const float* __restrict__ add, bf16* __restrict__ y16, int ldy,
                                                        WinGeom g, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * 8, wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n4 = C >> 2, np4 = ld >> 2, ny4 = ldy >> 2;
  for (long long r0 = wid * RW; r0 < rows; r0 += nwarps * RW) {
    long long m[RW];
    float4 v[RW][NJ];
    // lane k works out the map row of the group's row k (the index arithmetic is a dozen integer divisions: once per row, not per lane)
    const long long rl = r0 + (lane < RW ? lane : 0);
    const int rc = (int)(rl < rows ? rl : rows - 1);                   // a short last group repeats the last row (same values written twice)
    const int ml = (MODE == 1 || MODE == 2) ? win_src_row(g, rc) : rc;
#pragma unroll
    for (int k = 0; k < RW; ++k) {
      m[k] = __shfl_sync(0xffffffffu, ml, k);
      const float4* xr = reinterpret_cast<const float4*>(x + (m[k] < 0 ? 0 : m[k]) * ld);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int i = lane + 32 * j;
        v[k][j] = i < np4 ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (MODE == 2) {
#pragma unroll
      for (int k = 0; k < RW; ++k) {
        if (r0 + k >= rows || m[k] < 0) continue;                    // the repeated row must not be added twice; padding tokens are dropped
        const float4* ar = reinterpret_cast<const float4*>(add + (r0 + k) * ld);
        float4* xr = reinterpret_cast<float4*>(x + m[k] * ld);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int i = lane + 32 * j;
          if (i < np4) {
            const float4 a = __ldg(ar + i);
            v[k][j] = make_float4(v[k][j].x + a.x, v[k][j].y + a.y, v[k][j].z + a.z, v[k][j].w + a.w);
            xr[i] = v[k][j];                                         // pad columns: 0 + 0
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RW; ++k) {
      if (r0 + k >= rows) continue;
      if (m[k] < 0) {                                                // a padding token: a zero row in window order (MODE 1), nothing otherwise
        if (MODE == 1) {
          uint2* yz = reinterpret_cast<uint2*>(y16 + (r0 + k) * ldy);
          for (int i = lane; i < ny4; i += 32) yz[i] = make_uint2(0u, 0u);
        }
        continue;
      }
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (lane + 32 * j < n4) s += v[k][j].x + v[k][j].y + v[k][j].z + v[k][j].w;
      const float mean = warp_sum(s) / (float)C;
      float m2 = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (lane + 32 * j < n4) {
          const float a = v[k][j].x - mean, bb = v[k][j].y - mean, c = v[k][j].z - mean, d = v[k][j].w - mean;
          m2 += a * a + bb * bb + c * c + d * d;
        }
      const float rstd = rsqrtf(warp_sum(m2) / (float)C + eps);
      uint2* yr = MODE == 3 ? nullptr : reinterpret_cast<uint2*>(y16 + (MODE == 1 ? r0 + k : m[k]) * ldy);
      float4* xo = reinterpret_cast<float4*>(x + m[k] * ld);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int i = lane + 32 * j;
        if (i >= (MODE == 3 ? np4 : ny4)) break;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n4) {
          const float4 gw = __ldg(reinterpret_cast<const float4*>(w) + i), gb = __ldg(reinterpret_cast<const float4*>(b) + i);   // L1-resident
          o = make_float4((v[k][j].x - mean) * rstd * gw.x + gb.x, (v[k][j].y - mean) * rstd * gw.y + gb.y,
                          (v[k][j].z - mean) * rstd * gw.z + gb.z, (v[k][j].w - mean) * rstd * gw.w + gb.w);
        }
        if (MODE == 3) xo[i] = o;
        else yr[i] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
}
template <int MODE>
static void ln_window(float* x, int ld, int C, const float* w, const float* b, float eps, const float* add, bf16* y16, const WinGeom& g,
                      long long rows, cudaStream_t st) {
  const int nj = (ld / 4 + 31) / 32;
  VG_CHECK(ld % 4 == 0 && nj <= 8 && rows < (1LL << 31), "ln_window: rows of at most 1024 channels, fewer than 2^31 rows");
  auto grid = [&](int rw) { return (unsigned)std::min<long long>((rows + 8LL * rw - 1) / (8LL * rw), 148 * 4); };
  if (nj == 1) ln_window_kernel<MODE, 1, 4><<<grid(4), 256, 0, st>>>(x, ld, C, w, b, eps, add, y16, ld, g, rows);
  else if (nj == 2) ln_window_kernel<MODE, 2, 4><<<grid(4), 256, 0, st>>>(x, ld, C, w, b, eps, add, y16, ld, g, rows);
  else if (nj <= 4) ln_window_kernel<MODE, 4, 2><<<grid(2), 256, 0, st>>>(x, ld, C, w, b, eps, add, y16, ld, g, rows);
  else ln_window_kernel<MODE, 8, 1><<<grid(1), 256, 0, st>>>(x, ld, C, w, b, eps, add, y16, ld, g, rows);
  VG_CUDA(cudaGetLastError());
}

// PatchEmbed3D operand (patch (1,4,4)): frames NCHW fp32 [n, 3, R, R] → A [n * (R/4)^2, 64] bf16, column c*16 + dy*4 + dx (the
// order of proj.weight.reshape(96, 48)), columns 48..63 zero
__global__ void __launch_bounds__(256) patch_im2col_kernel(const float* __restrict__ fr, bf16* __restrict__ A, int R, long long n16) {
  const int G = R >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
    const long long tok = i >> 4;
    const int j = (int)(i & 15);
    uint2 o = make_uint2(0u, 0u);
    if (j < 12) {
      const int c = j >> 2, dy = j & 3;
      const long long n = tok / (G * G);
      const int rem = (int)(tok - n * G * G), py = rem / G, px = rem - py * G;
      const float4 v = __ldg(reinterpret_cast<const float4*>(fr + ((n * 3 + c) * R + py * 4 + dy) * (long long)R + px * 4));
      o = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
    reinterpret_cast<uint2*>(A)[i] = o;
  }
}

// PatchMerging (:291-308), even H and W: out[(b, d, h2, w2), k*C + c] = LayerNorm_4C(cat(x[2h2, 2w2], x[2h2+1, 2w2], x[2h2, 2w2+1],
// x[2h2+1, 2w2+1])) as bf16.  A warp walks output tokens with a grid stride, RW tokens (4 source rows of NJ float4 per lane each) in
// registers at a time: one read of the map.
template <int NJ, int RW>
__global__ void __launch_bounds__(256, 4) merge_ln_kernel(const float* __restrict__ x, int ld, int C, int H, int W, const float* __restrict__ w,
                                                          const float* __restrict__ b, float eps, bf16* __restrict__ out, long long tokens) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * 8, wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int H2 = H >> 1, W2 = W >> 1, n4 = C >> 2;
  for (long long t0 = wid * RW; t0 < tokens; t0 += nwarps * RW) {
    float4 v[RW][4][NJ];
#pragma unroll
    for (int q = 0; q < RW; ++q) {
      const long long t = t0 + q < tokens ? t0 + q : tokens - 1;
      const long long bd = t / (H2 * W2);
      const int rem = (int)(t - bd * H2 * W2), h2 = rem / W2, w2 = rem - h2 * W2;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4* src = reinterpret_cast<const float4*>(x + ((bd * H + 2 * h2 + (k & 1)) * W + 2 * w2 + (k >> 1)) * (long long)ld);
#pragma unroll
        for (int j = 0; j < NJ; ++j) v[q][k][j] = lane + 32 * j < n4 ? __ldg(src + lane + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int q = 0; q < RW; ++q) {
      if (t0 + q >= tokens) continue;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < NJ; ++j) s += v[q][k][j].x + v[q][k][j].y + v[q][k][j].z + v[q][k][j].w;      // lanes beyond the row hold zeros
      const float mean = warp_sum(s) / (float)(4 * C);
      float m2 = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (lane + 32 * j < n4) {
            const float a = v[q][k][j].x - mean, bb = v[q][k][j].y - mean, c = v[q][k][j].z - mean, d = v[q][k][j].w - mean;
            m2 += a * a + bb * bb + c * c + d * d;
          }
      const float rstd = rsqrtf(warp_sum(m2) / (float)(4 * C) + eps);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int i = lane + 32 * j;
          if (i < n4) {
            const float4 ww = __ldg(reinterpret_cast<const float4*>(w + k * C) + i), bv = __ldg(reinterpret_cast<const float4*>(b + k * C) + i);
            const float4 a = v[q][k][j];
            reinterpret_cast<uint2*>(out + (t0 + q) * 4 * C + k * C)[i] =
                make_uint2(pack_bf16((a.x - mean) * rstd * ww.x + bv.x, (a.y - mean) * rstd * ww.y + bv.y),
                           pack_bf16((a.z - mean) * rstd * ww.z + bv.z, (a.w - mean) * rstd * ww.w + bv.w));
          }
        }
    }
  }
}
static void merge_ln(const float* x, int ld, int C, int H, int W, const float* w, const float* b, float eps, bf16* out, long long tokens,
                     cudaStream_t st) {
  const int nj = (C / 4 + 31) / 32;
  VG_CHECK(C % 4 == 0 && nj <= 3, "merge_ln: at most 384 channels before PatchMerging");
  auto grid = [&](int rw) { return (unsigned)std::min<long long>((tokens + 8LL * rw - 1) / (8LL * rw), 148 * 4); };
  if (nj == 1) merge_ln_kernel<1, 2><<<grid(2), 256, 0, st>>>(x, ld, C, H, W, w, b, eps, out, tokens);
  else if (nj == 2) merge_ln_kernel<2, 1><<<grid(1), 256, 0, st>>>(x, ld, C, H, W, w, b, eps, out, tokens);
  else merge_ln_kernel<3, 1><<<grid(1), 256, 0, st>>>(x, ld, C, H, W, w, b, eps, out, tokens);
  VG_CUDA(cudaGetLastError());
}
// strided rows fp32 [rows, ld] → compact bf16 / fp32 [rows, C]
__global__ void __launch_bounds__(256) rows_out_kernel(const float* __restrict__ x, int ld, int C, bf16* __restrict__ o16, float* __restrict__ o32,
                                                       long long n4) {
  const int c4 = C >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4;
    const int j = (int)(i - r * c4);
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * ld) + j);
    if (o16 != nullptr) reinterpret_cast<uint2*>(o16)[i] = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
    if (o32 != nullptr) reinterpret_cast<float4*>(o32)[i] = a;
  }
}

static int grid_for(long long n) { return (int)std::min<long long>((n + 255) / 256, 148 * 32); }
static int pad64(int c) { return (c + 63) / 64 * 64; }

// ------------------------------------------------------------------------------------------------ weights
void SwinNet::pack(const HasFn& has, const GetFn& get, const std::function<bf16*(const float*, size_t)>& to_bf16,
                   const std::function<float*(const float*, size_t)>& to_f32) {
  const int nb = (2 * wd - 1) * (2 * wh - 1) * (2 * ww - 1), N = wd * wh * ww;
  // relative position index (video_swin_transformer.py:97-112)
  std::vector<int> idx((size_t)N * N);
  for (int a = 0; a < N; ++a)
    for (int b = 0; b < N; ++b) {
      const int da = a / (wh * ww), ha = (a / ww) % wh, wa = a % ww, db = b / (wh * ww), hb = (b / ww) % wh, wb = b % ww;
      idx[(size_t)a * N + b] = (da - db + wd - 1) * (2 * wh - 1) * (2 * ww - 1) + (ha - hb + wh - 1) * (2 * ww - 1) + (wa - wb + ww - 1);
    }
  // W [o, in] → channel-padded [op, inp] bf16 (zeros in the pad rows / columns); bias [o] → [op]
  auto lin_pad = [&](const std::string& name, int o, int in, int op, int inp, bf16*& W, float** bias) {
    const float* w = get(name + ".weight", {o, in});
    std::vector<float> wp((size_t)op * inp, 0.f);
    for (int r = 0; r < o; ++r) std::copy(w + (size_t)r * in, w + (size_t)(r + 1) * in, wp.begin() + (size_t)r * inp);
    W = to_bf16(wp.data(), wp.size());
    if (bias != nullptr) {
      const float* bb = get(name + ".bias", {o});
      std::vector<float> bp(op, 0.f);
      std::copy(bb, bb + o, bp.begin());
      *bias = to_f32(bp.data(), bp.size());
    }
  };
  for (int s = 0; s < kStages; ++s) {
    Stage& S = st[s];
    S.C = embed << s; S.Cp = pad64(S.C); S.Nqkv = (3 * S.C + 127) / 128 * 128; S.heads = heads[s];   // q|k|v width padded to 128-column GEMM tiles
    const std::string lp = "vid.layers." + std::to_string(s) + ".";
    if (!has(lp + "blocks.0.attn.qkv.weight")) continue;
    const int C = S.C, Cp = S.Cp;
    const float inv_scale = std::sqrt((float)(C / S.heads));
    S.blocks.resize(depths[s]);
    for (int i = 0; i < depths[s]; ++i) {
      const std::string p = lp + "blocks." + std::to_string(i) + ".";
      Block& k = S.blocks[i];
      k.n1w = to_f32(get(p + "norm1.weight", {C}), C); k.n1b = to_f32(get(p + "norm1.bias", {C}), C);
      k.n2w = to_f32(get(p + "norm2.weight", {C}), C); k.n2b = to_f32(get(p + "norm2.bias", {C}), C);
      lin_pad(p + "attn.qkv", 3 * C, C, S.Nqkv, Cp, k.Wqkv, &k.bqkv);
      lin_pad(p + "attn.proj", C, C, Cp, Cp, k.Wproj, &k.bproj);
      lin_pad(p + "mlp.fc1", 4 * C, C, 4 * C, Cp, k.Wfc1, &k.bfc1);
      lin_pad(p + "mlp.fc2", C, 4 * C, Cp, 4 * C, k.Wfc2, &k.bfc2);
      const float* tab = get(p + "attn.relative_position_bias_table", {nb, S.heads});
      std::vector<float> sb((size_t)S.heads * N * N);
      for (int h = 0; h < S.heads; ++h)
        for (size_t ab = 0; ab < (size_t)N * N; ++ab) sb[(size_t)h * N * N + ab] = tab[(size_t)idx[ab] * S.heads + h] * inv_scale;
      k.sbias = to_bf16(sb.data(), sb.size());
    }
    S.loaded = true;
    const std::string dp = "vid.downsamples." + std::to_string(s) + ".";
    if (s + 1 < kStages && has(dp + "reduction.weight")) {
      S.mnw = to_f32(get(dp + "norm.weight", {4 * C}), 4 * C); S.mnb = to_f32(get(dp + "norm.bias", {4 * C}), 4 * C);
      lin_pad(dp + "reduction", 2 * C, 4 * C, pad64(2 * C), 4 * C, S.Wred, nullptr);
    }
  }
  if (has("vid.patch_embed.proj.weight")) {
    const float* w = get("vid.patch_embed.proj.weight", {embed, 3, 1, 4, 4});
    std::vector<float> wp((size_t)pad64(embed) * 64, 0.f);
    for (int r = 0; r < embed; ++r) std::copy(w + (size_t)r * 48, w + (size_t)(r + 1) * 48, wp.begin() + (size_t)r * 64);
    Wpe = to_bf16(wp.data(), wp.size());
    std::vector<float> bp(pad64(embed), 0.f);
    const float* bb = get("vid.patch_embed.proj.bias", {embed});
    std::copy(bb, bb + embed, bp.begin());
    bpe = to_f32(bp.data(), bp.size());
    pnw = to_f32(get("vid.patch_embed.norm.weight", {embed}), embed); pnb = to_f32(get("vid.patch_embed.norm.bias", {embed}), embed);
    full = st[0].loaded && st[1].loaded && st[2].loaded && st[3].loaded && st[0].Wred && st[1].Wred && st[2].Wred;
    VG_CHECK(full, "Video-Swin: 'vid.patch_embed' is given but a stage or a downsample layer is missing");
  }
}

// ------------------------------------------------------------------------------------------------ workspace
// units: the largest (rows incl. window padding) x (padded channels) of any stage; qkv_units: the same for the packed q|k|v rows
void SwinNet::ensure_workspace(size_t units, size_t qkv_units, size_t frame_rows) {
  if (units <= cap_units && qkv_units <= cap_qkv && frame_rows <= cap_frames) return;
  units = std::max(units, cap_units); qkv_units = std::max(qkv_units, cap_qkv); frame_rows = std::max(frame_rows, cap_frames);
  release();
  const size_t u = units;
  VG_CUDA(cudaMalloc(&x32, u * 4)); VG_CUDA(cudaMalloc(&y32, u * 4));
  VG_CUDA(cudaMalloc(&xn, u * 2)); VG_CUDA(cudaMalloc(&xw, u * 2)); VG_CUDA(cudaMalloc(&ao, u * 2)); VG_CUDA(cudaMalloc(&xm, u * 2));
  VG_CUDA(cudaMalloc(&qkv, qkv_units * 2 + 4096)); VG_CUDA(cudaMalloc(&hid, u * 4 * 2));
  if (frame_rows) VG_CUDA(cudaMalloc(&a0, frame_rows * 64 * 2));
  cap_units = u; cap_qkv = qkv_units; cap_frames = frame_rows;
}
// workspace demand of stage s on a (clips, D, H, W) map
void SwinNet::stage_units(int s, int clips, int D, int H, int W, size_t& units, size_t& qkv_units) const {
  const size_t Dp = (D + wd - 1) / wd * wd, Hp = (H + wh - 1) / wh * wh, Wp = (W + ww - 1) / ww * ww;
  const size_t rows_p = (size_t)clips * Dp * Hp * Wp;
  units = std::max(units, rows_p * st[s].Cp);
  qkv_units = std::max(qkv_units, rows_p * st[s].Nqkv);
}

void SwinNet::release() {
  for (void* q : {(void*)x32, (void*)y32, (void*)xn, (void*)xw, (void*)qkv, (void*)ao, (void*)hid, (void*)xm, (void*)a0})
    if (q) cudaFree(q);
  x32 = y32 = nullptr; xn = xw = qkv = ao = hid = xm = a0 = nullptr; cap_units = cap_qkv = cap_frames = 0;
  for (auto& m : masks) { if (m.second.rid) cudaFree(m.second.rid); if (m.second.gset) cudaFree(m.second.gset); }
  masks.clear();
}

// ------------------------------------------------------------------------------------------------ one stage
void SwinNet::run_stage(int s, int clips, int D, int H, int W, cudaStream_t st_) {
  Stage& S = st[s];
  VG_CHECK(S.loaded, "Video-Swin: the weights of stage " + std::to_string(s) + " ('vid.layers." + std::to_string(s) + ".*') were not given");
  const int C = S.C, Cp = S.Cp;
  const int w_d = std::min(wd, D), w_h = std::min(wh, H), w_w = std::min(ww, W);     // get_window_size (:53-66)
  VG_CHECK(w_d == wd && w_h == wh && w_w == ww, "Video-Swin: maps smaller than the window (8 frames, 7x7 positions) are not supported");
  // sides that are not multiples of the window are zero-padded at their end for the attention half (:205-211,233-234)
  const int Dp = (D + wd - 1) / wd * wd, Hp = (H + wh - 1) / wh * wh, Wp = (W + ww - 1) / ww * ww;
  const bool padded = Dp != D || Hp != H || Wp != W;
  const int shd = D > wd ? wd / 2 : 0, shh = H > wh ? wh / 2 : 0, shw = W > ww ? ww / 2 : 0;
  const int N = wd * wh * ww, nW = (Dp / wd) * (Hp / wh) * (Wp / ww), groups = clips * nW;
  const long long rows = (long long)clips * D * H * W, rows_p = (long long)clips * Dp * Hp * Wp;
  const bool contiguous = H == wh && W == ww && !padded;          // a window = a run of whole frames: the partition is the identity
  const float scale = 1.0f / std::sqrt((float)(C / S.heads));
  // region ids of the shifted windows (compute_mask, :311-325), built once per map shape
  MaskTab* mt = nullptr;
  if (shd || shh || shw) {
    auto key = std::make_tuple(Dp, Hp, Wp, clips);
    MaskTab& m = masks[key];
    if (m.rid == nullptr) {
      auto region = [](int x, int X, int w, int sft) { return sft == 0 ? 0 : (x < X - w ? 0 : (x < X - sft ? 1 : 2)); };
      std::vector<std::vector<uint8_t>> sets(1, std::vector<uint8_t>(N, 0));
      std::vector<uint8_t> gs(nW);
      for (int g = 0; g < nW; ++g) {
        const int wwi = g % (Wp / ww), hwi = (g / (Wp / ww)) % (Hp / wh), dwi = g / ((Wp / ww) * (Hp / wh));
        std::vector<uint8_t> v(N);
        for (int n = 0; n < N; ++n) {
          const int dd = n / (wh * ww), hh = (n / ww) % wh, wl = n % ww;
          v[n] = (uint8_t)(region(dwi * wd + dd, Dp, wd, shd) * 9 + region(hwi * wh + hh, Hp, wh, shh) * 3 + region(wwi * ww + wl, Wp, ww, shw));
        }
        const bool constant = std::all_of(v.begin(), v.end(), [&](uint8_t q) { return q == v[0]; });
        int id = 0;
        if (!constant) {
          auto it = std::find(sets.begin() + 1, sets.end(), v);
          id = (int)(it - sets.begin());
          if (it == sets.end()) sets.push_back(v);
        }
        gs[g] = (uint8_t)id;
      }
      m.h_rid.clear();
      for (auto& v : sets) { v.resize(512, 0); m.h_rid.insert(m.h_rid.end(), v.begin(), v.end()); }   // rows of 512 bytes
      m.h_gset.resize((size_t)groups);
      for (int c = 0; c < clips; ++c) std::copy(gs.begin(), gs.end(), m.h_gset.begin() + (size_t)c * nW);
      VG_CUDA(cudaMalloc(&m.rid, m.h_rid.size()));
      VG_CUDA(cudaMalloc(&m.gset, m.h_gset.size()));
      VG_CUDA(cudaMemcpyAsync(m.rid, m.h_rid.data(), m.h_rid.size(), cudaMemcpyHostToDevice, st_));
      VG_CUDA(cudaMemcpyAsync(m.gset, m.h_gset.data(), m.h_gset.size(), cudaMemcpyHostToDevice, st_));
      m.nW = nW;
    }
    mt = &m;
  }
  if (Cp != C) VG_CUDA(cudaMemsetAsync(ao, 0, (size_t)rows_p * Cp * 2, st_));   // pad columns of the attention output stay zero
  for (size_t i = 0; i < S.blocks.size(); ++i) {
    Block& k = S.blocks[i];
    const bool shifted = (i % 2 == 1) && (shd || shh || shw);
    const WinGeom g{clips, D, H, W, Dp, Hp, Wp, wd, wh, ww, shifted ? shd : 0, shifted ? shh : 0, shifted ? shw : 0};
    const bool gather = shifted || !contiguous;
    const long long rows_a = gather ? rows_p : rows;   // rows of the attention half (window order, padding tokens included)
    const bf16* a = xn;
    if (gather) {   // LN1 + window partition + cyclic roll in one pass over the rows
      ln_window<1>(x32, Cp, C, k.n1w, k.n1b, 1e-5f, nullptr, xw, g, rows_a, st_);
      a = xw;
    } else {
      ln_window<0>(x32, Cp, C, k.n1w, k.n1b, 1e-5f, nullptr, xn, g, rows, st_);
    }
    { GemmEpi ep; ep.C = qkv; ep.ldc = S.Nqkv; ep.bias = k.bqkv; ep.bias_ld = S.Nqkv;
      gemm_bf16_tn(a, Cp, k.Wqkv, Cp, (int)rows_a, S.Nqkv, Cp, ep, st_); }
    window_attn_tc(qkv, S.Nqkv, ao, Cp, groups, N, S.heads, k.sbias, shifted ? mt->rid : nullptr, shifted ? mt->gset : nullptr, scale, st_);
    if (gather) {   // proj → fp32 in window order; window reverse + un-roll + residual add + LN2 in one pass over the rows
      GemmEpi ep; ep.C = y32; ep.ldc = Cp; ep.c_f32 = 1; ep.bias = k.bproj; ep.bias_ld = Cp;
      gemm_bf16_tn(ao, Cp, k.Wproj, Cp, (int)rows_a, Cp, Cp, ep, st_);
      ln_window<2>(x32, Cp, C, k.n2w, k.n2b, 1e-5f, y32, xn, g, rows_a, st_);
    } else {
      GemmEpi ep; ep.C = y32; ep.ldc = Cp; ep.c_f32 = 1; ep.bias = k.bproj; ep.bias_ld = Cp; ep.res32 = x32; ep.ldres32 = Cp;
      gemm_bf16_tn(ao, Cp, k.Wproj, Cp, (int)rows, Cp, Cp, ep, st_);
      std::swap(x32, y32);
      ln_window<0>(x32, Cp, C, k.n2w, k.n2b, 1e-5f, nullptr, xn, g, rows, st_);
    }
    { GemmEpi ep; ep.C = hid; ep.ldc = 4 * C; ep.bias = k.bfc1; ep.bias_ld = 4 * C; ep.act = ACT_GELU;
      gemm_bf16_tn(xn, Cp, k.Wfc1, Cp, (int)rows, 4 * C, Cp, ep, st_); }
    { GemmEpi ep; ep.C = y32; ep.ldc = Cp; ep.c_f32 = 1; ep.bias = k.bfc2; ep.bias_ld = Cp; ep.res32 = x32; ep.ldres32 = Cp;
      gemm_bf16_tn(hid, 4 * C, k.Wfc2, 4 * C, (int)rows, Cp, 4 * C, ep, st_); }
    std::swap(x32, y32);
    launches += 7;
  }
  VG_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ entry points
int SwinNet::forward_stage4(const float* x, int clips, int D, int H, int W, bf16* out_bf16, float* out_f32, cudaStream_t st_) {
  Stage& S = st[3];
  VG_CHECK(S.loaded, "vgqa_swin_stage needs the 'vid.layers.3.blocks.*' weights");
  VG_CHECK(clips >= 1 && D >= 1, "vgqa_swin_stage: bad shape");
  const int launches0 = launches;
  const long long rows = (long long)clips * D * H * W;
  { size_t u = 0, q = 0; stage_units(3, clips, D, H, W, u, q); ensure_workspace(u, q, 0); }
  VG_CUDA(cudaMemcpyAsync(x32, x, (size_t)rows * S.C * 4, cudaMemcpyDeviceToDevice, st_));
  run_stage(3, clips, D, H, W, st_);
  const long long n4 = rows * (S.C / 4);
  rows_out_kernel<<<grid_for(n4), 256, 0, st_>>>(x32, S.Cp, S.C, out_bf16, out_f32, n4);
  ++launches;
  VG_CUDA(cudaGetLastError());
  return launches - launches0;
}

int SwinNet::forward_full(const float* frames, int clips, int T, int R, bf16* out_bf16, float* out_f32, float* const* stage_out,
                          cudaStream_t st_) {
  VG_CHECK(full, "vgqa_swin_backbone needs the whole 'vid.*' state dict (patch_embed, layers.0-3, downsamples.0-2)");
  VG_CHECK(clips >= 1 && T >= wd && R >= 32 * wh && R % 32 == 0,
           "vgqa_swin_backbone: at least 8 frames per clip and a frame side that is a multiple of 32, at least 224");
  const int launches0 = launches;
  int H = R / 4;
  long long rows = (long long)clips * T * H * H;
  {
    size_t u = 0, q = 0;
    for (int s = 0, h = H; s < kStages; ++s, h /= 2) stage_units(s, clips, T, h, h, u, q);
    ensure_workspace(u, q, (size_t)rows);
  }
  {  // PatchEmbed3D (:426-443): 4x4 patches → GEMM [rows, 64] x [128, 64]^T + bias → LayerNorm(96), in place
    const long long n16 = rows * 16;
    patch_im2col_kernel<<<grid_for(n16), 256, 0, st_>>>(frames, a0, R, n16);
    GemmEpi ep; ep.C = x32; ep.ldc = st[0].Cp; ep.c_f32 = 1; ep.bias = bpe; ep.bias_ld = st[0].Cp;
    gemm_bf16_tn(a0, 64, Wpe, 64, (int)rows, st[0].Cp, 64, ep, st_);
    ln_window<3>(x32, st[0].Cp, st[0].C, pnw, pnb, 1e-5f, nullptr, nullptr, WinGeom{}, rows, st_);
    launches += 3;
  }
  for (int s = 0; s < kStages; ++s) {
    Stage& S = st[s];
    run_stage(s, clips, T, H, H, st_);
    if (stage_out != nullptr && stage_out[s] != nullptr) {
      const long long n4 = rows * (S.C / 4);
      rows_out_kernel<<<grid_for(n4), 256, 0, st_>>>(x32, S.Cp, S.C, nullptr, stage_out[s], n4);
      ++launches;
    }
    if (s + 1 < kStages) {   // PatchMerging (:291-308): 2x2 gather + LayerNorm(4C) → reduction GEMM → the next stage's stream
      VG_CHECK(H % 2 == 0, "Video-Swin: odd map side before PatchMerging");
      const long long tokens = rows / 4;
      merge_ln(x32, S.Cp, S.C, H, H, S.mnw, S.mnb, 1e-5f, xm, tokens, st_);
      GemmEpi ep; ep.C = y32; ep.ldc = st[s + 1].Cp; ep.c_f32 = 1;
      gemm_bf16_tn(xm, 4 * S.C, S.Wred, 4 * S.C, (int)tokens, st[s + 1].Cp, 4 * S.C, ep, st_);
      std::swap(x32, y32);
      launches += 2;
      H /= 2;
      rows = tokens;
    }
  }
  const long long n4 = rows * (st[3].C / 4);
  rows_out_kernel<<<grid_for(n4), 256, 0, st_>>>(x32, st[3].Cp, st[3].C, out_bf16, out_f32, n4);
  ++launches;
  VG_CUDA(cudaGetLastError());
  return launches - launches0;
}

}  // namespace vg
