// Attention kernels of the grounding hot path (head_dim = 32, 8 heads, d_model = 256).
//
//  mha32_kernel  — standard multi-head attention over small groups: per-frame encoder self-attention
//                  (S = 2P+L tokens; reference vgqa/core/decoder/modal_encoder.py:172 →
//                  F.multi_head_attention_forward), the decoders' temporal self-attention across T frames
//                  (query_decoder.py:294,469) and TemporalSampling's T x L cross-attention
//                  (language/bert_module.py:59-80).  Flash-style: online softmax over 64-key chunks, exp2
//                  with the scale folded in, fp32 statistics.
//  xattn1_kernel — "one query per frame" multi-head cross-attention with the key/value projections
//                  ABSORBED into the query/output side: scores_h(m) = q~_h · mem_m (+ q2_h · kpos_h(m)),
//                  ctx_h = sum_m p_h(m) mem_m.  Replaces TimeDecoder cross_attn_image (query_decoder.py:472-480),
//                  PosDecoder's conditional cross-attention (query_decoder.py:305-369 + attention.py:116-260)
//                  and SpatialActivation's BertAttention_Cross (classifier.py:73-78).  The memory-side
//                  Linear layers (92% of the reference decoder FLOPs) are never executed: mem is read once
//                  per layer at HBM speed.  Arithmetic intensity ~16 FLOP/B (HBM-bound), so warp-level
//                  HMMA (mma.sync m16n8k16) is used rather than tcgen05.
#include <stdlib.h>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vg {

static constexpr float kLog2e = 1.4426950408889634f;

// ------------------------------------------------------------------------------------------------
// mha32
// ------------------------------------------------------------------------------------------------
struct Mha32Params {
  const bf16* Q; const bf16* K; const bf16* V; bf16* O;
  const uint8_t* kmask;  // [groups, Sk] (1 = padded key) or nullptr
  int ldq, ldk, ldv, ldo;
  int Sq, Sk;            // rows per group
  float scale_log2e;     // softmax scale * log2(e)
};

// 64-byte rows (32 bf16), 16-byte chunks XOR-swizzled by (row>>1)&3 → conflict-free ldmatrix.
__device__ __forceinline__ uint32_t swz64(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

__global__ void __launch_bounds__(128) mha32_kernel(const Mha32Params p) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int g = blockIdx.x >> 3, h = blockIdx.x & 7;  // the 8 heads of a group run together → K/V lines shared in L2
  const int Sq = p.Sq, Sk = p.Sk;
  const int Sq_pad = (Sq + 15) & ~15, Sk_pad = (Sk + 63) & ~63;
  uint8_t* sQ = sm;
  uint8_t* sK = sQ + Sq_pad * 64;
  uint8_t* sV = sK + Sk_pad * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- stage Q, K, V head slices (rows of 64 B) into shared memory; zero the padding rows
  const bf16* gq = p.Q + (size_t)g * Sq * p.ldq + h * 32;
  const bf16* gk = p.K + (size_t)g * Sk * p.ldk + h * 32;
  const bf16* gv = p.V + (size_t)g * Sk * p.ldv + h * 32;
  for (int i = tid; i < Sq_pad * 4; i += 128) {
    const int r = i >> 2, c = i & 3;
    if (r < Sq) cp_async16(sQ + swz64(r, c), gq + (size_t)r * p.ldq + c * 8);
    else *reinterpret_cast<uint4*>(sQ + swz64(r, c)) = make_uint4(0, 0, 0, 0);
  }
  for (int i = tid; i < Sk_pad * 4; i += 128) {
    const int r = i >> 2, c = i & 3;
    if (r < Sk) {
      cp_async16(sK + swz64(r, c), gk + (size_t)r * p.ldk + c * 8);
      cp_async16(sV + swz64(r, c), gv + (size_t)r * p.ldv + c * 8);
    } else {
      *reinterpret_cast<uint4*>(sK + swz64(r, c)) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sV + swz64(r, c)) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const uint32_t sQa = smem_u32(sQ), sKa = smem_u32(sK), sVa = smem_u32(sV);
  const uint8_t* km = p.kmask ? p.kmask + (size_t)g * Sk : nullptr;
  const int gid = lane >> 2, tq = lane & 3;

  for (int qb = warp; qb * 16 < Sq; qb += 4) {
    // Q fragments for the two k-steps (dims 0-15, 16-31)
    uint32_t qf[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int row = qb * 16 + (lane & 15);
      const int chunk = ks * 2 + (lane >> 4);
      ldmatrix_x4(qf[ks], sQa + swz64(row, chunk));
    }
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kc = 0; kc < Sk_pad; kc += 64) {
      float s[8][4];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
        // B fragments: K rows (keys) x 32 dims; matrices: (dims 0-7),(8-15),(16-23),(24-31)
        uint32_t kf[4];
        const int row = kc + nb * 8 + (lane & 7);
        const int chunk = lane >> 3;
        ldmatrix_x4(kf, sKa + swz64(row, chunk));
        uint32_t b0[2] = {kf[0], kf[1]}, b1[2] = {kf[2], kf[3]};
        mma_16816(s[nb], qf[0], b0);
        mma_16816(s[nb], qf[1], b1);
      }
      // mask + scale (log2 domain)
      float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int key = kc + nb * 8 + tq * 2 + j;
          const bool dead = key >= Sk || (km != nullptr && km[key] != 0);
          s[nb][j] = dead ? -INFINITY : s[nb][j] * p.scale_log2e;
          s[nb][2 + j] = dead ? -INFINITY : s[nb][2 + j] * p.scale_log2e;
          cm0 = fmaxf(cm0, s[nb][j]);
          cm1 = fmaxf(cm1, s[nb][2 + j]);
        }
      }
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
      const float nm0 = fmaxf(m0, cm0), nm1 = fmaxf(m1, cm1);
      const float base0 = nm0 == -INFINITY ? 0.f : nm0, base1 = nm1 == -INFINITY ? 0.f : nm1;
      const float a0 = exp2f(m0 - base0), a1 = exp2f(m1 - base1);  // m == -inf → 0
      l0 *= a0; l1 *= a1;
#pragma unroll
      for (int i = 0; i < 4; ++i) { o[i][0] *= a0; o[i][1] *= a0; o[i][2] *= a1; o[i][3] *= a1; }
      m0 = nm0; m1 = nm1;
      uint32_t pf[4][4];  // A fragments of P for 4 k-steps of 16 keys
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const float p00 = exp2f(s[nb][0] - base0), p01 = exp2f(s[nb][1] - base0);
        const float p10 = exp2f(s[nb][2] - base1), p11 = exp2f(s[nb][3] - base1);
        l0 += p00 + p01; l1 += p10 + p11;
        pf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(p00, p01);
        pf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(p10, p11);
      }
      // O += P V : B[k=key][n=dim] from V rows via transposed ldmatrix
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 2; ++np) {  // pairs of 8-dim n-blocks
          uint32_t vf[4];
          const int row = kc + ks * 16 + (lane & 15);
          const int chunk = np * 2 + (lane >> 4);
          ldmatrix_x4_trans(vf, sVa + swz64(row, chunk));
          uint32_t b0[2] = {vf[0], vf[1]}, b1[2] = {vf[2], vf[3]};
          mma_16816(o[np * 2 + 0], pf[ks], b0);
          mma_16816(o[np * 2 + 1], pf[ks], b1);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = qb * 16 + gid, r1 = r0 + 8;
    bf16* go = p.O + (size_t)g * Sq * p.ldo + h * 32;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      if (r0 < Sq)
        *reinterpret_cast<uint32_t*>(go + (size_t)r0 * p.ldo + nb * 8 + tq * 2) = pack_bf16(o[nb][0] * i0, o[nb][1] * i0);
      if (r1 < Sq)
        *reinterpret_cast<uint32_t*>(go + (size_t)r1 * p.ldo + nb * 8 + tq * 2) = pack_bf16(o[nb][2] * i1, o[nb][3] * i1);
    }
  }
}

void mha32(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O, int ldo, int groups,
           int Sq, int Sk, const uint8_t* kmask, float scale, cudaStream_t stream) {
  VG_CHECK(groups > 0 && Sq > 0 && Sk > 0, "mha32: empty problem");
  Mha32Params p{Q, K, V, O, kmask, ldq, ldk, ldv, ldo, Sq, Sk, scale * kLog2e};
  const int Sq_pad = (Sq + 15) & ~15, Sk_pad = (Sk + 63) & ~63;
  const int smem = (Sq_pad + 2 * Sk_pad) * 64;
  VG_CHECK(smem <= 200 * 1024, "mha32: sequence too long for the shared-memory resident kernel");
  static int configured = 0;
  if (smem > configured) {
    VG_CUDA(cudaFuncSetAttribute(mha32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  mha32_kernel<<<groups * 8, 128, smem, stream>>>(p);
  VG_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// xattn1: one query per frame, absorbed K/V projections
// ------------------------------------------------------------------------------------------------
struct Xattn1Params {
  const bf16* qt;     // [F, 8*256] absorbed queries
  const bf16* mem;    // token 0 of frame 0; frame f token m at mem + (f*frame_stride + m)*256
  const bf16* posk;   // optional [*, Mk, 256] added to the keys (scores only); frame f uses posk + f*posk_fstride
  const bf16* q2;     // optional [F, 256]: per-head 32-d query for the kpos term
  const bf16* kpos;   // optional [*, Mk, ldkpos]; frame f uses kpos + f*kpos_fstride
  const uint8_t* kmask;  // optional [F, ldmask]
  const float* sbias; // optional [F, 8, ldsb] additive score term (unscaled), e.g. q~·pos or the kpos term as a GEMM
  int ldsb;
  bf16* ctx;          // [F, 8*256]
  float* att;         // optional [F, Mk]: minmax(sigmoid(sum_h p_h))  (classifier.py:75-78)
  long long frame_stride;  // in rows of 256
  long long posk_fstride, kpos_fstride;  // in elements
  int Mk, ldkpos, ldmask;
  float scale_log2e;
};

static constexpr int XROW = 264;  // padded smem row (elements): 528 B → conflict-free ldmatrix

static constexpr int XNT = 256;   // 8 warps

__global__ void __launch_bounds__(XNT) xattn1_kernel(const Xattn1Params p) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int f = blockIdx.x;
  const int Mk = p.Mk, Mp = (Mk + 15) & ~15;
  bf16* sMem = reinterpret_cast<bf16*>(sm);                       // [Mp][XROW]
  bf16* sQ = sMem + (size_t)Mp * XROW;                            // [8][XROW]
  bf16* sP = sQ + 8 * XROW;                                       // [16][Mp + 8]
  float* sS = reinterpret_cast<float*>(sP + 16 * (Mp + 8));       // [2 K-halves][8][Mp]
  bf16* sQ2 = reinterpret_cast<bf16*>(sS + 16 * Mp);              // [256]
  float* sRed = reinterpret_cast<float*>(sQ2 + 256);              // [16]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bf16* gmem = p.mem + (size_t)f * p.frame_stride * 256;
  const bf16* gpos = p.posk ? p.posk + (size_t)f * p.posk_fstride : nullptr;

  // ---- fill: keys (mem [+ pos]) , absorbed queries, q2; zero pad rows and P rows 8..15
  for (int i = tid; i < Mp * 32; i += XNT) {
    const int r = i >> 5, c = i & 31;
    bf16* dst = sMem + (size_t)r * XROW + c * 8;
    if (r < Mk) {
      if (gpos == nullptr) {
        cp_async16(dst, gmem + (size_t)r * 256 + c * 8);
      } else {
        const uint4 a = *reinterpret_cast<const uint4*>(gmem + (size_t)r * 256 + c * 8);
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(gpos + (size_t)r * 256 + c * 8));
        float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
        float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
        *reinterpret_cast<uint4*>(dst) =
            make_uint4(pack_bf16(a0.x + b0.x, a0.y + b0.y), pack_bf16(a1.x + b1.x, a1.y + b1.y),
                       pack_bf16(a2.x + b2.x, a2.y + b2.y), pack_bf16(a3.x + b3.x, a3.y + b3.y));
      }
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  }
  {
    const int r = tid >> 5, c = tid & 31;  // 8 rows x 32 chunks = 256 threads
    cp_async16(sQ + r * XROW + c * 8, p.qt + (size_t)f * 2048 + r * 256 + c * 8);
  }
  if (p.q2 != nullptr && tid < 32) cp_async16(sQ2 + tid * 8, p.q2 + (size_t)f * 256 + tid * 8);
  for (int i = tid; i < 8 * (Mp + 8) / 2; i += XNT) reinterpret_cast<uint32_t*>(sP + 8 * (Mp + 8))[i] = 0u;
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const uint32_t sMemA = smem_u32(sMem), sQA = smem_u32(sQ), sPA = smem_u32(sP);
  const int gid = lane >> 2, tq = lane & 3;

  // ---- phase 1: scores[key][head] = mem[key,:] · q~[head,:].  Work item = (16-key block, K half of 128 channels);
  //      two interleaved accumulators keep the HMMA dependency chains short.
  {
    const int kh = warp & 1;
    uint32_t bq[8][2];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      // B[k=channel][n=head]: q~ is [head][channel] → non-transposed ldmatrix, matrices (ch 0-7),(ch 8-15)
      const int row = lane & 7, half = (lane >> 3) & 1;
      ldmatrix_x2(bq[ks], sQA + (row * XROW + kh * 128 + ks * 16 + half * 8) * 2);
    }
    for (int kb = warp >> 1; kb * 16 < Mp; kb += 4) {
      float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
      const int row = kb * 16 + (lane & 15);
      const uint32_t abase = sMemA + (row * XROW + kh * 128 + (lane >> 4) * 8) * 2;
#pragma unroll
      for (int ks = 0; ks < 8; ks += 2) {
        uint32_t a0[4], a1[4];
        ldmatrix_x4(a0, abase + ks * 32);
        ldmatrix_x4(a1, abase + (ks + 1) * 32);
        mma_16816(d0, a0, bq[ks]);
        mma_16816(d1, a1, bq[ks + 1]);
      }
      float* pl = sS + kh * 8 * Mp;
      const int k0 = kb * 16 + gid, k1 = k0 + 8, h0 = tq * 2;
      pl[h0 * Mp + k0] = d0[0] + d1[0]; pl[(h0 + 1) * Mp + k0] = d0[1] + d1[1];
      pl[h0 * Mp + k1] = d0[2] + d1[2]; pl[(h0 + 1) * Mp + k1] = d0[3] + d1[3];
    }
  }
  __syncthreads();
  // ---- optional second score term: q2_h · kpos_h(m)  (32-d per head, SIMT; kpos is L1/L2 resident)
  if (p.kpos != nullptr) {
    const bf16* gk = p.kpos + (size_t)f * p.kpos_fstride;
    for (int i = tid; i < Mk * 8; i += XNT) {
      const int m = i >> 3, h = i & 7;
      const uint4* kp = reinterpret_cast<const uint4*>(gk + (size_t)m * p.ldkpos + h * 32);
      const uint4* qp = reinterpret_cast<const uint4*>(sQ2 + h * 32);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 kv = __ldg(kp + c), qv = qp[c];
        float2 k0 = unpack_bf16(kv.x), k1 = unpack_bf16(kv.y), k2 = unpack_bf16(kv.z), k3 = unpack_bf16(kv.w);
        float2 q0 = unpack_bf16(qv.x), q1 = unpack_bf16(qv.y), q2 = unpack_bf16(qv.z), q3 = unpack_bf16(qv.w);
        acc += k0.x * q0.x + k0.y * q0.y + k1.x * q1.x + k1.y * q1.y + k2.x * q2.x + k2.y * q2.y + k3.x * q3.x + k3.y * q3.y;
      }
      sS[h * Mp + m] += acc;
    }
    __syncthreads();
  }
  // the keys buffer held mem+pos: re-stage the plain memory rows for the value side (L2 hit)
  if (gpos != nullptr) {
    for (int i = tid; i < Mk * 32; i += XNT) {
      const int r = i >> 5, c = i & 31;
      cp_async16(sMem + (size_t)r * XROW + c * 8, gmem + (size_t)r * 256 + c * 8);
    }
    cp_async_commit();
  }
  // ---- phase 2: softmax over keys, one warp per head
  const uint8_t* km = p.kmask ? p.kmask + (size_t)f * p.ldmask : nullptr;
  {
    const int h = warp;
    float* s0 = sS + h * Mp;
    const float* s1 = sS + (8 + h) * Mp;
    const float* sb = p.sbias ? p.sbias + ((size_t)f * 8 + h) * p.ldsb : nullptr;
    float mx = -INFINITY;
    for (int m = lane; m < Mk; m += 32) {
      float v = s0[m] + s1[m];
      if (sb != nullptr) v += __ldg(sb + m);
      v *= p.scale_log2e;
      if (km != nullptr && km[m] != 0) v = -INFINITY;
      s0[m] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    if (mx == -INFINITY) mx = 0.f;
    float sum = 0.f;
    for (int m = lane; m < Mk; m += 32) {
      const float e = exp2f(s0[m] - mx);
      s0[m] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int m = lane; m < Mp; m += 32) {
      const float pr = m < Mk ? s0[m] * inv : 0.f;
      if (m < Mk) s0[m] = pr;
      sP[h * (Mp + 8) + m] = __float2bfloat16(pr);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // ---- optional attention map: minmax(sigmoid(sum over heads))
  if (p.att != nullptr) {
    float lmin = INFINITY, lmax = -INFINITY;
    float* amap = sS + 8 * Mp;  // the second score plane is free now
    for (int m = tid; m < Mk; m += XNT) {
      float a = 0.f;
#pragma unroll
      for (int h = 0; h < 8; ++h) a += sS[h * Mp + m];
      a = 1.f / (1.f + __expf(-a));
      amap[m] = a;
      lmin = fminf(lmin, a); lmax = fmaxf(lmax, a);
    }
    lmin = -warp_max(-lmin); lmax = warp_max(lmax);
    if (lane == 0) { sRed[warp] = lmin; sRed[8 + warp] = lmax; }
    __syncthreads();
    float amin = sRed[0], amax = sRed[8];
#pragma unroll
    for (int w = 1; w < 8; ++w) { amin = fminf(amin, sRed[w]); amax = fmaxf(amax, sRed[8 + w]); }
    const float inv = 1.f / (amax - amin + 1e-6f);
    for (int m = tid; m < Mk; m += XNT) p.att[(size_t)f * Mk + m] = (amap[m] - amin) * inv;
  }
  // ---- phase 3: ctx[head][channel] = sum_key P[head][key] mem[key][channel]; warp w → channels [32w, 32w+32)
  {
    float d[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    for (int ks = 0; ks * 16 < Mp; ++ks) {
      uint32_t a[4];
      ldmatrix_x4(a, sPA + ((lane & 15) * (Mp + 8) + ks * 16 + (lane >> 4) * 8) * 2);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t vf[4];
        const int row = ks * 16 + (lane & 15);
        const int col = warp * 32 + np * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(vf, sMemA + (row * XROW + col) * 2);
        uint32_t b0[2] = {vf[0], vf[1]}, b1[2] = {vf[2], vf[3]};
        mma_16816(d[np * 2 + 0], a, b0);
        mma_16816(d[np * 2 + 1], a, b1);
      }
    }
    bf16* gc = p.ctx + (size_t)f * 2048 + gid * 256 + warp * 32 + tq * 2;  // rows 0..7 = heads
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) *reinterpret_cast<uint32_t*>(gc + nb * 8) = pack_bf16(d[nb][0], d[nb][1]);
  }
}

void xattn1(const bf16* qt, const bf16* mem, long long frame_stride_rows, int F, int Mk, const bf16* posk,
            long long posk_fstride, const bf16* q2, const bf16* kpos, int ldkpos, long long kpos_fstride,
            const uint8_t* kmask, int ldmask, float scale, bf16* ctx, float* att, cudaStream_t stream,
            const float* sbias, int ldsb) {
  VG_CHECK(F > 0 && Mk > 0, "xattn1: empty problem");
  {  // frame-invariant positional terms arrive as `sbias`: the streaming kernel (xattn_stream.cu) handles everything else
    static int use_stream = -1;
    if (use_stream < 0) { const char* e = getenv("VGQA_XATTN_STREAM"); use_stream = (e == nullptr || e[0] != '0') ? 1 : 0; }
    if (use_stream && posk == nullptr && q2 == nullptr && kpos == nullptr && xattn_stream_supported(Mk, frame_stride_rows)) {
      xattn_stream(qt, mem, frame_stride_rows, F, Mk, sbias, ldsb, kmask, ldmask, scale, ctx, att, stream);
      return;
    }
  }
  Xattn1Params p;
  p.qt = qt; p.mem = mem; p.posk = posk; p.q2 = q2; p.kpos = kpos; p.kmask = kmask; p.ctx = ctx; p.att = att;
  p.sbias = sbias; p.ldsb = ldsb;
  p.frame_stride = frame_stride_rows; p.posk_fstride = posk_fstride; p.kpos_fstride = kpos_fstride;
  p.Mk = Mk; p.ldkpos = ldkpos; p.ldmask = ldmask; p.scale_log2e = scale * kLog2e;
  const int Mp = (Mk + 15) & ~15;
  const int smem = (Mp * XROW + 8 * XROW + 16 * (Mp + 8) + 256) * 2 + (16 * Mp + 16) * 4;
  VG_CHECK(smem <= 220 * 1024, "xattn1: too many keys for the shared-memory resident kernel");
  static int configured = 0;
  if (smem > configured) {
    VG_CUDA(cudaFuncSetAttribute(xattn1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  xattn1_kernel<<<F, XNT, smem, stream>>>(p);
  VG_CUDA(cudaGetLastError());
}

}  // namespace vg
