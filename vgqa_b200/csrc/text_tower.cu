// Small kernels of the RoBERTa text tower (vgqa/core/language/bert.py:49,66-69: `RobertaModel(...)(**tokenized).last_hidden_state`;
// arithmetic = transformers' RobertaEmbeddings + RobertaLayer, see DESIGN.md "Text tower").  A query is
// L <= 64 tokens, so everything except the four Linear layers per block (tcgen05 GEMMs, gemm_tc.cu) is latency-bound SIMT:
//   roberta_embed_ln   word + position (cumsum of non-pad ids) + token-type embedding → LayerNorm          one warp per token
//   text_attn          softmax(q k^T / 8 + key padding mask) v, heads of 64, keys held in shared memory     one CTA per (query, head)
//   ln_rows_wide       LayerNorm over rows of any width % 32 == 0 (the GEMM epilogue has already added bias + residual)
#include "kernels.h"
#include "ptx.cuh"

namespace vg {

// x[r, :] = LN(word[ids[r]] + pos[pos_id(r)] + type0);  pos_id = ids[r] != pad ? pad + #(non-pad ids of the row's query up to r) : pad
__global__ void __launch_bounds__(256) roberta_embed_ln_kernel(const int* __restrict__ ids, const float* __restrict__ word,
                                                               const float* __restrict__ pos, const float* __restrict__ type0,
                                                               const float* __restrict__ w, const float* __restrict__ b,
                                                               float eps, float* __restrict__ x32, bf16* __restrict__ x, int rows,
                                                               int L, int Hd, int vocab, int max_pos, int pad_id) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int q = r / L, l = r - q * L;
  int cnt = 0;
  for (int i = lane; i <= l; i += 32) cnt += ids[q * L + i] != pad_id;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  int id = ids[r];
  const int pid = id != pad_id ? min(pad_id + cnt, max_pos - 1) : pad_id;
  id = min(max(id, 0), vocab - 1);
  const float* wr = word + (size_t)id * Hd;
  const float* pr = pos + (size_t)pid * Hd;
  float s = 0.f;
  for (int c = lane; c < Hd; c += 32) s += wr[c] + pr[c] + type0[c];
  const float mean = warp_sum(s) / (float)Hd;
  float m2 = 0.f;
  for (int c = lane; c < Hd; c += 32) { const float d = wr[c] + pr[c] + type0[c] - mean; m2 = fmaf(d, d, m2); }
  const float rstd = rsqrtf(warp_sum(m2) / (float)Hd + eps);
  for (int c = lane; c < Hd; c += 32) {
    const float y = (wr[c] + pr[c] + type0[c] - mean) * rstd * w[c] + b[c];
    x32[(size_t)r * Hd + c] = y;
    x[(size_t)r * Hd + c] = __float2bfloat16(y);
  }
}
void roberta_embed_ln(const int* ids, const float* word, const float* pos, const float* type0, const float* w, const float* b,
                      float eps, float* x32, bf16* x, int rows, int L, int Hd, int vocab, int max_pos, int pad_id, cudaStream_t st) {
  roberta_embed_ln_kernel<<<(rows + 7) / 8, 256, 0, st>>>(ids, word, pos, type0, w, b, eps, x32, x, rows, L, Hd, vocab, max_pos, pad_id);
  VG_CUDA(cudaGetLastError());
}

// y = LN(x) * w + b over rows of Hd <= 1024 fp32 values (one warp per row, the row held in registers: one global read);
// writes fp32 (optional; may alias x) and bf16
__global__ void __launch_bounds__(256) ln_rows_wide_kernel(const float* x, const float* __restrict__ w, const float* __restrict__ b,
                                                           float eps, float* y32, bf16* __restrict__ y, int rows, int Hd) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* xr = x + (size_t)r * Hd;
  float v[32];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < Hd ? xr[c] : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)Hd;
  float m2 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float d = lane + 32 * i < Hd ? v[i] - mean : 0.f;
    m2 = fmaf(d, d, m2);
  }
  const float rstd = rsqrtf(warp_sum(m2) / (float)Hd + eps);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    if (c < Hd) {
      const float o = (v[i] - mean) * rstd * w[c] + b[c];
      if (y32 != nullptr) y32[(size_t)r * Hd + c] = o;
      y[(size_t)r * Hd + c] = __float2bfloat16(o);
    }
  }
}
void ln_rows_wide(const float* x, const float* w, const float* b, float eps, float* y32, bf16* y, int rows, int Hd, cudaStream_t st) {
  VG_CHECK(Hd % 32 == 0 && Hd <= 1024, "ln_rows_wide: the row width must be a multiple of 32, at most 1024");
  ln_rows_wide_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, w, b, eps, y32, y, rows, Hd);
  VG_CUDA(cudaGetLastError());
}

// ctx[(q*L + l), h*64 + d] = sum_k softmax_k(scale * Q[l]·K[k] (+ -inf where pad[q, k])) V[k, d];  QKV rows = [q | k | v], each Hd wide
static constexpr int kTextMaxL = 64;
__global__ void __launch_bounds__(128) text_attn_kernel(const bf16* __restrict__ QKV, const uint8_t* __restrict__ pad,
                                                        bf16* __restrict__ ctx, int L, int Hd, float scale) {
  __shared__ float sk[kTextMaxL][65];   // +1: conflict-free column walks
  __shared__ float sv[kTextMaxL][64];
  __shared__ float sq[4][64];
  __shared__ float sp[4][kTextMaxL];
  const int q = blockIdx.x, h = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t ld = (size_t)3 * Hd;
  for (int i = threadIdx.x; i < L * 64; i += 128) {
    const int k = i >> 6, d = i & 63;
    const bf16* row = QKV + ((size_t)q * L + k) * ld + h * 64 + d;
    sk[k][d] = __bfloat162float(row[Hd]);
    sv[k][d] = __bfloat162float(row[2 * Hd]);
  }
  __syncthreads();
  for (int l = warp; l < L; l += 4) {
    const bf16* qr = QKV + ((size_t)q * L + l) * ld + h * 64;
    sq[warp][lane] = __bfloat162float(qr[lane]);
    sq[warp][lane + 32] = __bfloat162float(qr[lane + 32]);
    __syncwarp();
    float s[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int k = lane + 32 * t;
      float acc = -INFINITY;
      if (k < L && !(pad != nullptr && pad[(size_t)q * L + k] != 0)) {
        acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < 64; ++d) acc = fmaf(sq[warp][d], sk[k][d], acc);
        acc *= scale;
      }
      s[t] = acc;
    }
    const float mx = warp_max(fmaxf(s[0], s[1]));
    const float base = mx == -INFINITY ? 0.f : mx;
    const float p0 = __expf(s[0] - base), p1 = __expf(s[1] - base);
    const float inv = 1.f / warp_sum(p0 + p1);
    sp[warp][lane] = p0 * inv;
    sp[warp][lane + 32] = p1 * inv;
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int k = 0; k < L; ++k) {
      const float pk = sp[warp][k];
      o0 = fmaf(pk, sv[k][lane], o0);
      o1 = fmaf(pk, sv[k][lane + 32], o1);
    }
    bf16* out = ctx + ((size_t)q * L + l) * Hd + h * 64;
    out[lane] = __float2bfloat16(o0);
    out[lane + 32] = __float2bfloat16(o1);
    __syncwarp();
  }
}
void text_attn(const bf16* QKV, const uint8_t* pad, bf16* ctx, int Q, int L, int Hd, cudaStream_t st) {
  VG_CHECK(L >= 1 && L <= kTextMaxL, "text_attn: a query has at most 64 tokens");
  VG_CHECK(Hd % 64 == 0, "text_attn: the hidden size must be a multiple of 64 (heads of 64)");
  text_attn_kernel<<<dim3(Q, Hd / 64), 128, 0, st>>>(QKV, pad, ctx, L, Hd, 0.125f);
  VG_CUDA(cudaGetLastError());
}

}  // namespace vg
