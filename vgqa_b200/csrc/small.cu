// Bandwidth-bound kernels of the grounding hot path: layout conversion, LayerNorm + pooled means, tiny heads,
// on-device frame selection (no host sync), query seeding, box sine embedding and PostProcess.
// Coalesced 16-byte accesses, warp-shuffle reductions, fp32 statistics.
#include <algorithm>

#include "kernels.h"
#include "ptx.cuh"

namespace vg {

// ---------------------------------------------------------------- layout: NCHW fp32 → token-major bf16
// Replaces `flatten(2).permute(2,0,1)` + `torch.cat` of CrossModalEncoder.forward (modal_encoder.py:50-66).
__global__ void __launch_bounds__(256) nchw_to_tokens_kernel(const float* __restrict__ in, long long in_fstride,
                                                             bf16* __restrict__ X, float* __restrict__ X32,
                                                             const float* __restrict__ pos, long long pos_fstride,
                                                             bf16* __restrict__ XP, int S, int tok0, int P) {
  __shared__ float tile[32][33];
  const int f = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  __shared__ float ptile[32][33];
  const float* src = in + (size_t)f * in_fstride;
  const float* psrc = pos ? pos + (size_t)f * pos_fstride : nullptr;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, p = p0 + tx;
    tile[ty + 8 * k][tx] = p < P ? src[(size_t)c * P + p] : 0.f;
    if (psrc) ptile[ty + 8 * k][tx] = p < P ? psrc[(size_t)c * P + p] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int p = p0 + ty + 8 * k;
    if (p < P) {
      const size_t o = ((size_t)f * S + tok0 + p) * 256 + c0 + tx;
      X[o] = __float2bfloat16(tile[tx][ty + 8 * k]);
      if (X32 != nullptr) X32[o] = tile[tx][ty + 8 * k];
      if (XP != nullptr) XP[o] = __float2bfloat16(tile[tx][ty + 8 * k] + ptile[tx][ty + 8 * k]);
    }
  }
}
void nchw_to_tokens(const float* in, long long in_fstride, bf16* X, float* X32, const float* pos, long long pos_fstride,
                    bf16* XP, int F, int S, int tok0, int P, cudaStream_t st) {
  dim3 grid((P + 31) / 32, 8, F);
  nchw_to_tokens_kernel<<<grid, 256, 0, st>>>(in, in_fstride, X, X32, pos, pos_fstride, XP, S, tok0, P);
  VG_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- PositionEmbeddingSine(128, normalize=True)
// vgqa/core/vision/position_encoding.py:50-91: cumsum of ~mask along y / x, normalised by the last row / column (+1e-6) and
// scaled to 2*pi, divided by 10000^(2*floor(i/2)/128), sin on even / cos on odd feature indices, cat(pos_y, pos_x) → NCHW.
// One CTA per frame (a frame has at most a few hundred positions); mask == nullptr → nothing is padded.
__global__ void __launch_bounds__(256) pos_sine_kernel(const uint8_t* __restrict__ mask, float* __restrict__ out, int H, int W) {
  extern __shared__ float emb[];   // [2][H*W]: y_embed, x_embed (normalised, scaled)
  const int f = blockIdx.x, P = H * W;
  const uint8_t* m = mask ? mask + (size_t)f * P : nullptr;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const int y = p / W, x = p % W;
    float cy = 0.f, ty = 0.f, cx = 0.f, tx = 0.f;
    for (int yy = 0; yy < H; ++yy) { const float v = (m && m[yy * W + x]) ? 0.f : 1.f; ty += v; if (yy <= y) cy += v; }
    for (int xx = 0; xx < W; ++xx) { const float v = (m && m[y * W + xx]) ? 0.f : 1.f; tx += v; if (xx <= x) cx += v; }
    const float scale = 6.283185307179586f;
    emb[p] = cy / (ty + 1e-6f) * scale;
    emb[P + p] = cx / (tx + 1e-6f) * scale;
  }
  __syncthreads();
  float* o = out + (size_t)f * 256 * P;
  for (int i = threadIdx.x; i < 256 * P; i += blockDim.x) {
    const int c = i / P, p = i % P, k = c & 127;
    const float dim_t = powf(10000.f, (float)(2 * (k >> 1)) / 128.f);
    const float a = emb[(c >> 7) * P + p] / dim_t;
    o[i] = (k & 1) ? cosf(a) : sinf(a);
  }
}
void pos_sine(const uint8_t* mask, float* out, int frames, int H, int W, cudaStream_t st) {
  pos_sine_kernel<<<frames, 256, (size_t)2 * H * W * sizeof(float), st>>>(mask, out, H, W);
  VG_CUDA(cudaGetLastError());
}

// channels-last bf16 features (vgqa_inputs.feat_layout = 1): a (frame, position) IS a token row — copy, widen and add pos
__global__ void __launch_bounds__(256) rows_to_tokens_kernel(const bf16* __restrict__ in, const bf16* __restrict__ pos_tokens,
                                                             int pos_frames, bf16* __restrict__ X, float* __restrict__ X32,
                                                             bf16* __restrict__ XP, int S, int tok0, int P, long long n_chunks) {
  // one thread per 8 channels (16 bytes in, 16 + 32 + 16 bytes out)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i >> 5;            // f * P + p
    const int c = (int)(i & 31) * 8;
    const long long f = row / P;
    const int p = (int)(row - f * P);
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + row * 256 + c));
    const size_t o = ((size_t)f * S + tok0 + p) * 256 + c;
    *reinterpret_cast<uint4*>(X + o) = q;
    const float2 a = unpack_bf16(q.x), b = unpack_bf16(q.y), cc = unpack_bf16(q.z), d = unpack_bf16(q.w);
    if (X32 != nullptr) {
      *reinterpret_cast<float4*>(X32 + o) = make_float4(a.x, a.y, b.x, b.y);
      *reinterpret_cast<float4*>(X32 + o + 4) = make_float4(cc.x, cc.y, d.x, d.y);
    }
    if (XP != nullptr) {
      const size_t po = ((size_t)(pos_frames > 1 ? f : 0) * S + tok0 + p) * 256 + c;
      const uint4 pq = __ldg(reinterpret_cast<const uint4*>(pos_tokens + po));
      const float2 pa = unpack_bf16(pq.x), pb = unpack_bf16(pq.y), pc = unpack_bf16(pq.z), pd = unpack_bf16(pq.w);
      *reinterpret_cast<uint4*>(XP + o) = make_uint4(pack_bf16(a.x + pa.x, a.y + pa.y), pack_bf16(b.x + pb.x, b.y + pb.y),
                                                     pack_bf16(cc.x + pc.x, cc.y + pc.y), pack_bf16(d.x + pd.x, d.y + pd.y));
    }
  }
}
void rows_to_tokens(const bf16* in, const bf16* pos_tokens, int pos_frames, bf16* X, float* X32, bf16* XP, int F, int S, int tok0,
                    int P, cudaStream_t st) {
  const long long n = (long long)F * P * 32;
  const int grid = (int)std::min<long long>((n + 255) / 256, 148 * 16);
  rows_to_tokens_kernel<<<grid, 256, 0, st>>>(in, pos_tokens, pos_frames, X, X32, XP, S, tok0, P, n);
  VG_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256) text_to_tokens_kernel(const float* __restrict__ text, bf16* __restrict__ X,
                                                             float* __restrict__ X32, bf16* __restrict__ XP, int T, int S,
                                                             int tok0, int L) {
  const int f = blockIdx.x, b = f / T;
  for (int i = threadIdx.x; i < L * 64; i += 256) {
    const int l = i >> 6, c = (i & 63) * 4;
    uint2 o = make_uint2(0u, 0u);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (text != nullptr) {
      v = *reinterpret_cast<const float4*>(text + ((size_t)b * L + l) * 256 + c);
      o = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
    *reinterpret_cast<uint2*>(X + ((size_t)f * S + tok0 + l) * 256 + c) = o;
    if (X32 != nullptr) *reinterpret_cast<float4*>(X32 + ((size_t)f * S + tok0 + l) * 256 + c) = v;
    if (XP != nullptr) *reinterpret_cast<uint2*>(XP + ((size_t)f * S + tok0 + l) * 256 + c) = o;  // pos = 0 on text rows
  }
}
void text_to_tokens(const float* text, bf16* X, float* X32, bf16* XP, int F, int T, int S, int tok0, int L,
                    cudaStream_t st) {
  text_to_tokens_kernel<<<F, 256, 0, st>>>(text, X, X32, XP, T, S, tok0, L);
  VG_CUDA(cudaGetLastError());
}

// modal_encoder.py:46,53,60,65
__global__ void build_encoded_mask_kernel(const uint8_t* vis_mask, const uint8_t* text_mask, uint8_t* out, int T, int P,
                                          int L) {
  const int f = blockIdx.x, S = 2 * P + L;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    uint8_t v;
    if (s < P) v = (vis_mask && s != 0) ? vis_mask[(size_t)f * P + s] : 0;
    else if (s < P + L) v = text_mask ? text_mask[(size_t)(f / T) * L + (s - P)] : 0;
    else v = (vis_mask && s != P + L) ? vis_mask[(size_t)f * P + (s - P - L)] : 0;
    out[(size_t)f * S + s] = v;
  }
}
void build_encoded_mask(const uint8_t* vis_mask, const uint8_t* text_mask, uint8_t* out, int F, int T, int P, int L,
                        cudaStream_t st) {
  build_encoded_mask_kernel<<<F, 128, 0, st>>>(vis_mask, text_mask, out, T, P, L);
  VG_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- final encoder LayerNorm + pooled means
// SpatialTemporalEncoder.forward tail (modal_encoder.py:135-140) + adaptive_avg_pool2d of classifier.py:33.
// v[8] holds this lane's 8 channels of one 256-wide fp32 row; normalises in place (two-pass statistics)
__device__ __forceinline__ void ln_row8(const float* src, float (&v)[8], float eps, const float* w, const float* b,
                                        int c0) {
  const float4 x0 = *reinterpret_cast<const float4*>(src), x1 = *reinterpret_cast<const float4*>(src + 4);
  v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.f / 256.f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float t = v[i] - mean; q = fmaf(t, t, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / 256.f) + eps);
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c0)), w1 = __ldg(reinterpret_cast<const float4*>(w + c0 + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c0)), b1 = __ldg(reinterpret_cast<const float4*>(b + c0 + 4));
  v[0] = (v[0] - mean) * rstd * w0.x + b0.x; v[1] = (v[1] - mean) * rstd * w0.y + b0.y;
  v[2] = (v[2] - mean) * rstd * w0.z + b0.z; v[3] = (v[3] - mean) * rstd * w0.w + b0.w;
  v[4] = (v[4] - mean) * rstd * w1.x + b1.x; v[5] = (v[5] - mean) * rstd * w1.y + b1.y;
  v[6] = (v[6] - mean) * rstd * w1.z + b1.z; v[7] = (v[7] - mean) * rstd * w1.w + b1.w;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

__global__ void __launch_bounds__(256) enc_finalize_kernel(const float* __restrict__ X, const float* __restrict__ w,
                                                           const float* __restrict__ b, float eps, bf16* __restrict__ Xf,
                                                           float* __restrict__ frames_cls, bf16* __restrict__ pool_vis,
                                                           bf16* __restrict__ pool_vid, float* __restrict__ pool_vis32,
                                                           float* __restrict__ pool_vid32, int S, int P, int L) {
  __shared__ float red[3][8][256];
  const int f = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c0 = lane * 8;
  float aall[8], avis[8], avid[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) aall[i] = avis[i] = avid[i] = 0.f;
  for (int s = warp; s < S; s += 8) {
    const size_t off = ((size_t)f * S + s) * 256 + c0;
    float v[8];
    ln_row8(X + off, v, eps, w, b, c0);
    *reinterpret_cast<uint4*>(Xf + off) = pack8(v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      aall[i] += v[i];
      if (s < P) avis[i] += v[i];
      if (s >= P + L) avid[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[0][warp][c0 + i] = aall[i]; red[1][warp][c0 + i] = avis[i]; red[2][warp][c0 + i] = avid[i]; }
  __syncthreads();
  const int c = threadIdx.x;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { t0 += red[0][k][c]; t1 += red[1][k][c]; t2 += red[2][k][c]; }
  frames_cls[(size_t)f * 256 + c] = t0 / S;
  pool_vis[(size_t)f * 256 + c] = __float2bfloat16(t1 / P);
  pool_vid[(size_t)f * 256 + c] = __float2bfloat16(t2 / P);
  pool_vis32[(size_t)f * 256 + c] = t1 / P;
  pool_vid32[(size_t)f * 256 + c] = t2 / P;
}
void enc_finalize(const float* X, const float* w, const float* b, float eps, bf16* Xf, float* frames_cls, bf16* pool_vis,
                  bf16* pool_vid, float* pool_vis32, float* pool_vid32, int F, int S, int P, int L, cudaStream_t st) {
  enc_finalize_kernel<<<F, 256, 0, st>>>(X, w, b, eps, Xf, frames_cls, pool_vis, pool_vid, pool_vis32, pool_vid32, S, P, L);
  VG_CUDA(cudaGetLastError());
}

// grounding_net.py:119 (f_text_cls) and :131 (f_text_cls[:, :1] as the SpatialActivation init query).
// Two steps so that a frame-sharded clip can all-reduce the fp32 sums in between (vgqa_set_sharding).
__global__ void __launch_bounds__(256) text_sum_kernel(const bf16* __restrict__ Xf, float* __restrict__ sums, int T, int S,
                                                       int P, int L) {
  const int b = blockIdx.x / L, l = blockIdx.x % L, c = threadIdx.x;
  float acc = 0.f;
  for (int t = 0; t < T; ++t) acc += __bfloat162float(Xf[(((size_t)b * T + t) * S + P + l) * 256 + c]);
  sums[((size_t)b * L + l) * 256 + c] = acc;
}
void text_sum(const bf16* Xf, float* sums, int B, int T, int S, int P, int L, cudaStream_t st) {
  text_sum_kernel<<<B * L, 256, 0, st>>>(Xf, sums, T, S, P, L);
  VG_CUDA(cudaGetLastError());
}
__global__ void __launch_bounds__(256) text_finish_kernel(const float* __restrict__ sums, float inv_t, bf16* __restrict__ ftext,
                                                          bf16* __restrict__ q0, float* __restrict__ q0_32, int T, int L) {
  const int b = blockIdx.x / L, l = blockIdx.x % L, c = threadIdx.x;
  const float m32 = sums[((size_t)b * L + l) * 256 + c] * inv_t;
  const bf16 m = __float2bfloat16(m32);
  ftext[((size_t)b * L + l) * 256 + c] = m;
  if (l == 0)
    for (int t = 0; t < T; ++t) {
      q0[((size_t)b * T + t) * 256 + c] = m;
      q0_32[((size_t)b * T + t) * 256 + c] = m32;
    }
}
void text_finish(const float* sums, float inv_t_global, bf16* ftext, bf16* q0, float* q0_32, int B, int T, int L,
                 cudaStream_t st) {
  text_finish_kernel<<<B * L, 256, 0, st>>>(sums, inv_t_global, ftext, q0, q0_32, T, L);
  VG_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- tiny-N heads (vocab 1/20/34, sted 2, actioness 1, box 4)
__global__ void __launch_bounds__(256) rowvec_head_kernel(const bf16* __restrict__ x, int ldx, const float* __restrict__ w,
                                                          const float* __restrict__ b, float* __restrict__ y, int ldy,
                                                          int rows, int N, int act) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const uint4 raw = *reinterpret_cast<const uint4*>(x + (size_t)r * ldx + lane * 8);
  float2 a = unpack_bf16(raw.x), bq = unpack_bf16(raw.y), c = unpack_bf16(raw.z), d = unpack_bf16(raw.w);
  for (int j = 0; j < N; ++j) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (size_t)j * 256 + lane * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (size_t)j * 256 + lane * 8 + 4));
    float s = a.x * w0.x + a.y * w0.y + bq.x * w0.z + bq.y * w0.w + c.x * w1.x + c.y * w1.y + d.x * w1.z + d.y * w1.w;
    s = warp_sum(s);
    if (lane == 0) {
      s += b ? b[j] : 0.f;
      if (act == 1) s = 1.f / (1.f + expf(-s));
      y[(size_t)r * ldy + j] = s;
    }
  }
}
void rowvec_head(const bf16* x, int ldx, const float* w, const float* b, float* y, int ldy, int rows, int N, int act,
                 cudaStream_t st) {
  rowvec_head_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, ldx, w, b, y, ldy, rows, N, act);
  VG_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- frame selection on device
// grounding_net.py:125-128 (theta = 0.45, fallback to all frames) — replaces nonzero().tolist() host syncs.
__global__ void __launch_bounds__(256) select_pass1_kernel(const float* lfm, const float* lfa, float theta,
                                                           const float* force_w, float* att, float* w, float* K, int T) {
  __shared__ float cnt[8];
  const int b = blockIdx.x;
  float local = 0.f;
  for (int t = threadIdx.x; t < T; t += 256) {
    const size_t i = (size_t)b * T + t;
    const float a = 0.5f * (1.f / (1.f + expf(-lfm[i])) + 1.f / (1.f + expf(-lfa[i])));
    att[i] = a;
    const float sel = force_w ? force_w[i] : (a > theta ? 1.f : 0.f);
    w[i] = sel;
    local += sel;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) cnt[threadIdx.x >> 5] = local;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += cnt[k];
  if (threadIdx.x == 0) K[b] = tot;   // local count; the fall-back is applied by select_finish
}
void select_pass1(const float* lfm, const float* lfa, float theta, const float* force_w, float* att, float* w, float* K,
                  int B, int T, cudaStream_t st) {
  select_pass1_kernel<<<B, 256, 0, st>>>(lfm, lfa, theta, force_w, att, w, K, T);
  VG_CUDA(cudaGetLastError());
}
// grounding_net.py:143-150: sigmoid(actioness) > 0.5
__global__ void __launch_bounds__(256) select_pass2_kernel(const float* act_sig, const float* force_w, float* w,
                                                           float* K, int T) {
  __shared__ float cnt[8];
  const int b = blockIdx.x;
  float local = 0.f;
  for (int t = threadIdx.x; t < T; t += 256) {
    const size_t i = (size_t)b * T + t;
    const float sel = force_w ? force_w[i] : (act_sig[i] > 0.5f ? 1.f : 0.f);
    w[i] = sel;
    local += sel;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) cnt[threadIdx.x >> 5] = local;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += cnt[k];
  if (threadIdx.x == 0) K[b] = tot;
}
void select_pass2(const float* act_sig, const float* force_w, float* w, float* K, int B, int T, cudaStream_t st) {
  select_pass2_kernel<<<B, 256, 0, st>>>(act_sig, force_w, w, K, T);
  VG_CUDA(cudaGetLastError());
}

// `choose_index or nonzero(att_sequences > 0)` (grounding_net.py:128,150): if no frame of the clip was selected
// (K = count over ALL ranks of a sharded clip) every frame is used — the sigmoid average is always > 0.
__global__ void __launch_bounds__(256) select_finish_kernel(float* w, float* K, int T, int T_global) {
  const int b = blockIdx.x;
  if (K[b] != 0.f) return;
  for (int t = threadIdx.x; t < T; t += 256) w[(size_t)b * T + t] = 1.f;
  __syncthreads();
  if (threadIdx.x == 0) K[b] = (float)T_global;
}
void select_finish(float* w, float* K, int B, int T, int T_global, cudaStream_t st) {
  select_finish_kernel<<<B, 256, 0, st>>>(w, K, T, T_global);
  VG_CUDA(cudaGetLastError());
}

// classifier.py:80 `head(query).mean(0)` and grounding_net.py:135-136 `(enc[chosen] * att_map).mean((0, 1))` over the
// chosen frames, as local fp32 sums (red[b] = [64 logit sums | 256 seed sums]) + a finishing step; a sharded clip
// all-reduces `red` in between.
__global__ void __launch_bounds__(320) masked_sums_kernel(const float* __restrict__ logit_rows, int ldx, int N,
                                                          const float* __restrict__ part, const float* __restrict__ w,
                                                          float* __restrict__ red, int T) {
  const int b = blockIdx.x, j = threadIdx.x;
  float acc = 0.f;
  if (j < 64) {
    if (j < N)
      for (int t = 0; t < T; ++t) acc += w[(size_t)b * T + t] * logit_rows[((size_t)b * T + t) * ldx + j];
  } else {
    for (int t = 0; t < T; ++t) acc += part[((size_t)b * T + t) * 256 + (j - 64)];   // part is already weighted by w
  }
  red[(size_t)b * 320 + j] = acc;
}
void masked_sums(const float* logit_rows, int ldx, int N, const float* part, const float* w, float* red, int B, int T,
                 cudaStream_t st) {
  masked_sums_kernel<<<B, 320, 0, st>>>(logit_rows, ldx, N, part, w, red, T);
  VG_CUDA(cudaGetLastError());
}
__global__ void __launch_bounds__(320) seed_finish_kernel(const float* __restrict__ red, const float* __restrict__ K,
                                                          float* __restrict__ logits_r, int N, float* __restrict__ q,
                                                          bf16* __restrict__ tgt, int ldt, float* __restrict__ tgt32, int T,
                                                          int P) {
  const int b = blockIdx.x, j = threadIdx.x;
  const float v = red[(size_t)b * 320 + j];
  if (j < 64) {
    if (j < N) logits_r[(size_t)b * N + j] = v / K[b];
    return;
  }
  const int c = j - 64;
  const float acc = v / (K[b] * (float)P);
  q[(size_t)b * 256 + c] = acc;
  const bf16 h = __float2bfloat16(acc);
  for (int t = 0; t < T; ++t) {  // query_decoder.py:102,114 expand
    tgt[((size_t)b * T + t) * ldt + c] = h;
    tgt32[((size_t)b * T + t) * 256 + c] = acc;
  }
}
void seed_finish(const float* red, const float* K, float* logits_r, int N, float* q, bf16* tgt, int ldt, float* tgt32,
                 int B, int T, int P, cudaStream_t st) {
  seed_finish_kernel<<<B, 320, 0, st>>>(red, K, logits_r, N, q, tgt, ldt, tgt32, T, P);
  VG_CUDA(cudaGetLastError());
}

// grounding_net.py:135-136 / 155-160: per-frame part of (enc[chosen] * att_map[..., None]).sum over tokens
__global__ void __launch_bounds__(256) seed_partial_kernel(const bf16* __restrict__ Xf, const float* __restrict__ att,
                                                           const float* __restrict__ w, float* __restrict__ part, int S,
                                                           int tok0, int P) {
  const int f = blockIdx.x, c = threadIdx.x;
  float acc = 0.f;
  if (w[f] != 0.f) {
    const bf16* base = Xf + ((size_t)f * S + tok0) * 256 + c;
    for (int p = 0; p < P; ++p) acc = fmaf(att[(size_t)f * P + p], __bfloat162float(base[(size_t)p * 256]), acc);
  }
  part[(size_t)f * 256 + c] = acc;
}
void seed_partial(const bf16* Xf, const float* att, const float* w, float* part, int F, int S, int tok0, int P,
                  cudaStream_t st) {
  seed_partial_kernel<<<F, 256, 0, st>>>(Xf, att, w, part, S, tok0, P);
  VG_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- anchors
// query_decoder.py:53-59,92-94: pos_fc = BertLN(256) → Linear(256,4) → ReLU → BertLN(4); then sigmoid
__global__ void __launch_bounds__(256) pos_fc_boxes_kernel(const float* __restrict__ fc, const float* ln0w,
                                                           const float* ln0b, const float* W, const float* bias,
                                                           const float* ln4w, const float* ln4b, float* boxes, int F) {
  const int f = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (f >= F) return;
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = fc[(size_t)f * 256 + lane * 8 + i]; s += v[i]; }
  const float mean = warp_sum(s) * (1.f / 256.f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float t = v[i] - mean; q = fmaf(t, t, q); }
  const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / 256.f) + 1e-12f);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * ln0w[lane * 8 + i] + ln0b[lane * 8 + i];
  float y[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) d = fmaf(v[i], W[j * 256 + lane * 8 + i], d);
    y[j] = fmaxf(warp_sum(d) + bias[j], 0.f);
  }
  const float m4 = 0.25f * (y[0] + y[1] + y[2] + y[3]);
  float v4 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) v4 += (y[j] - m4) * (y[j] - m4);
  const float r4 = 1.f / sqrtf(0.25f * v4 + 1e-12f);
  if (lane < 4) {
    const float z = (y[lane] - m4) * r4 * ln4w[lane] + ln4b[lane];
    boxes[(size_t)f * 4 + lane] = 1.f / (1.f + expf(-z));
  }
}
void pos_fc_boxes(const float* frames_cls, const float* ln0w, const float* ln0b, const float* W, const float* b,
                  const float* ln4w, const float* ln4b, float* boxes, int F, cudaStream_t st) {
  pos_fc_boxes_kernel<<<(F + 7) / 8, 256, 0, st>>>(frames_cls, ln0w, ln0b, W, b, ln4w, ln4b, boxes, F);
  VG_CUDA(cudaGetLastError());
}

// model_utils.py:15-40: order (y, x, w, h), 128 features each, sin on even / cos on odd indices
__global__ void __launch_bounds__(256) sine_embed_kernel(const float* __restrict__ boxes, bf16* __restrict__ sine, int F) {
  const int i = blockIdx.x * 256 + threadIdx.x;  // one (frame, feature pair) per thread: F * 256 pairs
  if (i >= F * 256) return;
  const int f = i >> 8, j = i & 255, grp = j >> 6, k = j & 63;  // pair k of group grp
  const int coord = grp == 0 ? 1 : (grp == 1 ? 0 : grp);           // y, x, w, h
  const float dim_t = powf(10000.f, (float)(2 * k) / 128.f);
  const float a = boxes[(size_t)f * 4 + coord] * 6.283185307179586f / dim_t;
  *reinterpret_cast<uint32_t*>(sine + (size_t)f * 512 + grp * 128 + 2 * k) = pack_bf16(sinf(a), cosf(a));
}
void sine_embed(const float* boxes, bf16* sine, int F, cudaStream_t st) {
  sine_embed_kernel<<<F, 256, 0, st>>>(boxes, sine, F);
  VG_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256) ln_rows_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                                      const float* __restrict__ b, float eps, bf16* __restrict__ y,
                                                      int ldy, int rows) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  float v[8];
  ln_row8(x + (size_t)r * ldx + lane * 8, v, eps, w, b, lane * 8);
  *reinterpret_cast<uint4*>(y + (size_t)r * ldy + lane * 8) = pack8(v);
}
void ln_rows(const float* x, int ldx, const float* w, const float* b, float eps, bf16* y, int ldy, int rows,
             cudaStream_t st) {
  ln_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, ldx, w, b, eps, y, ldy, rows);
  VG_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256) copy_cols_kernel(const bf16* __restrict__ src, int lds, bf16* __restrict__ dst,
                                                        int ldd, int rows, int cols8) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (size_t)rows * cols8) return;
  const size_t r = i / cols8, c = (i % cols8) * 8;
  *reinterpret_cast<uint4*>(dst + r * ldd + c) = *reinterpret_cast<const uint4*>(src + r * lds + c);
}
void copy_cols_bf16(const bf16* src, int lds, bf16* dst, int ldd, int rows, int cols, cudaStream_t st) {
  const size_t n = (size_t)rows * (cols / 8);
  copy_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, lds, dst, ldd, rows, cols / 8);
  VG_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- PostProcess (postprocessor.py:14-50)
__global__ void __launch_bounds__(256) postprocess_kernel(const float* __restrict__ boxes, const float* __restrict__ sted,
                                                          const float* __restrict__ sizes_hw, float* __restrict__ boxes_px,
                                                          int* __restrict__ sted_idx, int T) {
  extern __shared__ float sh[];  // ls[T], le[T]
  float* ls = sh;
  float* le = sh + T;
  __shared__ float rv[256];
  __shared__ int ri[256];
  __shared__ float red[2][8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float img_h = sizes_hw[b * 2 + 0], img_w = sizes_hw[b * 2 + 1];
  for (int t = tid; t < T; t += 256) {
    const float* bx = boxes + ((size_t)b * T + t) * 4;
    const float cx = bx[0], cy = bx[1], w = bx[2], h = bx[3];
    float* o = boxes_px + ((size_t)b * T + t) * 4;
    o[0] = fmaxf((cx - 0.5f * w) * img_w, 0.f);
    o[1] = fmaxf((cy - 0.5f * h) * img_h, 0.f);
    o[2] = fmaxf((cx + 0.5f * w) * img_w, 0.f);
    o[3] = fmaxf((cy + 0.5f * h) * img_h, 0.f);
  }
  // log-softmax over the T frames of each column
  for (int col = 0; col < 2; ++col) {
    float mx = -INFINITY;
    for (int t = tid; t < T; t += 256) mx = fmaxf(mx, sted[((size_t)b * T + t) * 2 + col]);
    mx = warp_max(mx);
    if ((tid & 31) == 0) red[0][tid >> 5] = mx;
    __syncthreads();
    mx = red[0][0];
#pragma unroll
    for (int k = 1; k < 8; ++k) mx = fmaxf(mx, red[0][k]);
    float sum = 0.f;
    for (int t = tid; t < T; t += 256) sum += expf(sted[((size_t)b * T + t) * 2 + col] - mx);
    sum = warp_sum(sum);
    if ((tid & 31) == 0) red[1][tid >> 5] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += red[1][k];
    const float lse = mx + logf(sum);
    float* dst = col == 0 ? ls : le;
    for (int t = tid; t < T; t += 256) dst[t] = sted[((size_t)b * T + t) * 2 + col] - lse;
    __syncthreads();
  }
  // argmax over s < e of ls[s] + le[e]; ties → lowest flat index s*T+e (torch.max semantics)
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int s = tid; s < T - 1; s += 256) {
    for (int e = s + 1; e < T; ++e) {
      const float v = ls[s] + le[e];
      const int idx = s * T + e;
      if (v > best || (v == best && idx < besti)) { best = v; besti = idx; }
    }
  }
  rv[tid] = best; ri[tid] = besti;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) {
      const float v = rv[tid + off]; const int i = ri[tid + off];
      if (v > rv[tid] || (v == rv[tid] && i < ri[tid])) { rv[tid] = v; ri[tid] = i; }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const int idx = ri[0] == 0x7fffffff ? 0 : ri[0];
    sted_idx[b * 2 + 0] = idx / T;
    sted_idx[b * 2 + 1] = idx % T;
  }
}
void postprocess(const float* boxes, const float* sted, const float* sizes_hw, float* boxes_px, int* sted_idx, int B,
                 int T, cudaStream_t st) {
  postprocess_kernel<<<B, 256, 2 * T * sizeof(float), st>>>(boxes, sted, sizes_hw, boxes_px, sted_idx, T);
  VG_CUDA(cudaGetLastError());
}

// ---- operands of the "score bias as a GEMM" formulation of the decoders' positional terms (model.cu, run_decoders)
__global__ void build_kpos_blockdiag_kernel(const bf16* __restrict__ kposb, int ldk, bf16* __restrict__ out, int Mk, int Mpad) {
  // grid (8 * Mpad, layers), 32 threads x 8 bf16 = one 256-wide row
  const int r = blockIdx.x, l = blockIdx.y, h = r / Mpad, m = r % Mpad, c = threadIdx.x;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (m < Mk && (c >> 2) == h) v = *reinterpret_cast<const uint4*>(kposb + (size_t)m * ldk + l * 256 + c * 8);
  *reinterpret_cast<uint4*>(out + ((size_t)l * 8 * Mpad + r) * 256 + c * 8) = v;
}
void build_kpos_blockdiag(const bf16* kposb, int ldk, bf16* out, int layers, int Mk, int Mpad, cudaStream_t st) {
  build_kpos_blockdiag_kernel<<<dim3(8 * Mpad, layers), 32, 0, st>>>(kposb, ldk, out, Mk, Mpad);
  VG_CUDA(cudaGetLastError());
}
__global__ void pad_rows_bf16_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int rows) {
  const int r = blockIdx.x, c = threadIdx.x;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (r < rows) v = *reinterpret_cast<const uint4*>(src + (size_t)r * 256 + c * 8);
  *reinterpret_cast<uint4*>(dst + (size_t)r * 256 + c * 8) = v;
}
void pad_rows_bf16(const bf16* src, bf16* dst, int rows, int rows_pad, cudaStream_t st) {
  pad_rows_bf16_kernel<<<rows_pad, 32, 0, st>>>(src, dst, rows);
  VG_CUDA(cudaGetLastError());
}

}  // namespace vg
