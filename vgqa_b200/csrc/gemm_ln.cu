// tcgen05 GEMM with a fused residual + LayerNorm epilogue (N = 256 = one full row per accumulator lane):
//
//   v   = act(A[M,K] · W[256,K]^T + bias) + res32          (res32: fp32 residual stream, optional)
//   y   = LayerNorm_256(v) * ln_w + ln_b                   (eps = 1e-5 nn.LayerNorm / 1e-12 BertLayerNorm)
//   C   = bf16(y)        next GEMM's A operand
//   C32 = y              fp32 residual stream (optional)
//   C2  = bf16(y + add2[row % period])   "x + pos" operand of the next layer's Q/K projection (optional)
//
// Replaces `norm(x + dropout(sublayer(x)))` of every post-norm block: modal_encoder.py:173-177,
// query_decoder.py:295-296,368-374,470,480-485, bert_module.py:92-96,137-141,205-209.
//
// These GEMMs are HBM-bound (out-proj: 1.5 KB/row moved for 131 kFLOP; FFN linear2: 6.5 KB/row for 1 MFLOP), so the
// design goal is that every byte moves as a full 128-byte line through the async proxy:
//   warp 0   TMA producer (A 128x64 + W 256x64 blocks, 3-stage ring)      warp 1   tcgen05.mma issuer, 2 TMEM accumulators
//   warps 2..5 epilogue, one accumulator row per thread:
//     pass 1  residual chunks arrive by TMA (32 rows x 128 B, double-buffered per warp) → v → shifted moments,
//             v parked back in TMEM (tcgen05.st)
//     pass 2  normalise, write fp32 / bf16 / bf16+pos slabs (128B-swizzled) → TMA stores.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace vg {

// NSPLIT = 2: a CTA pair (cluster) shares one 128-row tile, each CTA owning 128 of the 256 output columns (half of W, so half
// the bytes per CTA and twice the CTAs for the decoders' M = clips*T <= 4096-row GEMMs with K = 2048); the two halves of a
// row exchange their LayerNorm moments through distributed shared memory.
static constexpr int LBM = 128, LBK = 64;
template <int NSPLIT>
struct LnCfg {
  static constexpr int LBN = 256 / NSPLIT;
  static constexpr int kStages = NSPLIT == 1 ? 3 : 4;
  static constexpr int kABytes = LBM * LBK * 2, kBBytes = LBN * LBK * 2, kStageBytes = kABytes + kBBytes;
  static constexpr int kStaging = 4 * 16384;
  // [2 tile parities][128 rows] (mean, M2) written by the peer CTA (NSPLIT = 2); [2][NSPLIT source ranks][128] for wider clusters
  static constexpr int kXch = NSPLIT == 1 ? 0 : (NSPLIT == 2 ? 2 * 128 * 8 : 2 * NSPLIT * 128 * 8);
  static constexpr int kSmem = kStages * kStageBytes + kStaging + kXch + 3 * LBN * 4 + 1024 + 256;
};

struct LnParams {
  const float* bias; const float* ln_w; const float* ln_b;
  const bf16* add2;
  int M, K, act, add2_period, has_res, has_c32, has_c2;
  float eps;
};

__device__ __forceinline__ float ln_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int NSPLIT>
__global__ void __cluster_dims__(NSPLIT, 1, 1) __launch_bounds__(192, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_res, const __grid_constant__ CUtensorMap tma_c,
               const __grid_constant__ CUtensorMap tma_c32, const __grid_constant__ CUtensorMap tma_c2, const LnParams p) {
  using Cfg = LnCfg<NSPLIT>;
  constexpr int LBN = Cfg::LBN, kLnStages = Cfg::kStages, kLnABytes = Cfg::kABytes, kLnBBytes = Cfg::kBBytes;
  constexpr int kLnStageBytes = Cfg::kStageBytes, kLnStaging = Cfg::kStaging;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kLnStages * kLnABytes;
  uint8_t* smem_stg = smem + kLnStages * kLnStageBytes;
  float2* xch = reinterpret_cast<float2*>(smem_stg + kLnStaging);
  float* sbias = reinterpret_cast<float*>(smem_stg + kLnStaging + Cfg::kXch);
  float* slnw = sbias + LBN;
  float* slnb = slnw + LBN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(slnb + LBN);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kLnStages;
  uint64_t* tfull_bar = empty_bar + kLnStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;  // [4 warps][2]
  uint64_t* xbar = res_bar + 8;        // [2]: the peers' epilogue threads have delivered their moments (NSPLIT = 2 uses [0] only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.M + LBM - 1) / LBM;
  const int num_k = p.K / LBK;
  const int rank = NSPLIT == 1 ? 0 : (int)cluster_ctarank();
  const int col0 = rank * LBN;                              // first output column of this CTA
  const int tile0 = (int)blockIdx.x / NSPLIT, tile_step = (int)gridDim.x / NSPLIT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_b); tma_prefetch_desc(&tma_c);
    for (int s = 0; s < kLnStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
    for (int i = 0; i < 8; ++i) mbar_init(&res_bar[i], 1);
    mbar_init(&xbar[0], 128 * (NSPLIT > 1 ? NSPLIT - 1 : 1));
    mbar_init(&xbar[1], 128 * (NSPLIT > 1 ? NSPLIT - 1 : 1));
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * LBN);
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < LBN; i += 128) {
      sbias[i] = p.bias ? p.bias[col0 + i] : 0.f;
      slnw[i] = p.ln_w[col0 + i];
      slnb[i] = p.ln_b[col0 + i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (NSPLIT > 1) cluster_sync_all();   // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // PDL: the weight blocks of the first ring pass do not depend on the previous kernel — fetch them before waiting for it
      const int pre = tile0 < num_m ? (num_k < kLnStages ? num_k : kLnStages) : 0;
      for (int kb = 0; kb < pre; ++kb) {
        mbar_expect_tx(&full_bar[kb], kLnStageBytes);
        tma_load_2d(smem_b + kb * kLnBBytes, &tma_b, &full_bar[kb], kb * LBK, col0);
      }
      pdl_wait();
      pdl_launch_dependents();
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int mt = tile0; mt < num_m; mt += tile_step) {
        for (int kb = 0; kb < num_k; ++kb) {
          if (first && kb < pre) {
            tma_load_2d(smem_a + stage * kLnABytes, &tma_a, &full_bar[stage], kb * LBK, mt * LBM);
          } else {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], kLnStageBytes);
            tma_load_2d(smem_a + stage * kLnABytes, &tma_a, &full_bar[stage], kb * LBK, mt * LBM);
            tma_load_2d(smem_b + stage * kLnBBytes, &tma_b, &full_bar[stage], kb * LBK, col0);
          }
          if (++stage == kLnStages) { stage = 0; phase ^= 1; }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(LBM, LBN);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int mt = tile0; mt < num_m; mt += tile_step) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * LBN;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem_a + stage * kLnABytes));
          const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smem_b + stage * kLnBBytes));
#pragma unroll
          for (int k = 0; k < LBK / 16; ++k)
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == num_k - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == kLnStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5), thread = accumulator row =====================
    pdl_wait();   // the residual rows and the output buffers belong to earlier kernels of the stream
    const int quad = warp & 3;
    uint8_t* stg = smem_stg + (warp - 2) * 16384;
    uint8_t* bufA = stg;            // residual chunk (even) / bf16 output slab
    uint8_t* bufB = stg + 4096;     // residual chunk (odd)  / bf16 (y + pos) slab
    uint8_t* bufC = stg + 8192;     // fp32 output slab, columns [0,32) of the unit
    uint8_t* bufD = stg + 12288;    // fp32 output slab, columns [32,64) of the unit
    uint64_t* rbar = res_bar + (warp - 2) * 2;
    uint32_t rph0 = 0, rph1 = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int swz = lane & 7;
    uint32_t xph = 0;
    const uint32_t peer_xch = NSPLIT == 1 ? 0u : mapa_u32(smem_u32(xch), (uint32_t)(rank ^ 1));
    const uint32_t peer_xbar = NSPLIT == 1 ? 0u : mapa_u32(smem_u32(xbar), (uint32_t)(rank ^ 1));
    int tile_i = 0;
    for (int mt = tile0; mt < num_m; mt += tile_step, ++tile_i) {
      const int row0 = mt * LBM + quad * 32;
      const int row = row0 + lane;
      const bool valid = row < p.M;
      if (lane == 0) tma_store_wait_read<0>();  // the previous tile's stores have drained this warp's slabs
      __syncwarp();
      if (p.has_res && lane == 0) {
        mbar_expect_tx(&rbar[0], 4096); tma_load_2d(bufA, &tma_res, &rbar[0], col0, row0);
        mbar_expect_tx(&rbar[1], 4096); tma_load_2d(bufB, &tma_res, &rbar[1], col0 + 32, row0);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * LBN;
      uint32_t raw[32];
      float v[32];
      // ---------------- pass 1: v = act(acc + bias) + residual; moments; park v in TMEM
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < LBN / 32; ++c) {
        tmem_ld32(taddr + c * 32, raw);
        tmem_ld_wait();
        const float* sb = sbias + c * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) + sb[i];
        if (p.act == ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        } else if (p.act == ACT_GELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = ln_gelu(v[i]);
        }
        if (p.has_res) {
          uint8_t* rb = (c & 1) ? bufB : bufA;
          if (c & 1) { mbar_wait(&rbar[1], rph1); rph1 ^= 1; } else { mbar_wait(&rbar[0], rph0); rph0 ^= 1; }
          const uint8_t* rowp = rb + lane * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 r = *reinterpret_cast<const float4*>(rowp + ((i ^ swz) << 4));
            v[4 * i + 0] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
          }
          __syncwarp();  // every lane has consumed this buffer
          if (c + 2 < LBN / 32 && lane == 0) {
            mbar_expect_tx(&rbar[c & 1], 4096);
            tma_load_2d(rb, &tma_res, &rbar[c & 1], col0 + (c + 2) * 32, row0);
          }
        }
        if (c == 0) shift = v[0];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = v[i] - shift;
          s1 += d;
          s2 = fmaf(d, d, s2);
          raw[i] = __float_as_uint(v[i]);
        }
        tmem_st32(taddr + c * 32, raw);
      }
      tmem_st_wait();
      const float dm = s1 * (1.0f / LBN);
      float mean = shift + dm;
      float var = fmaxf(s2 * (1.0f / LBN) - dm * dm, 0.f);
      if (NSPLIT > 2) {
        // (mean, M2) of this CTA's columns → every peer's exchange slot [parity][my rank][row]; then the NSPLIT partial moments of
        // equal weight are combined: mean = avg(mean_q), M2 = sum M2_q + LBN * sum (mean_q - mean)^2.  One barrier per tile parity.
        const int par = tile_i & 1, rowi = quad * 32 + lane;
        const float m2 = var * (float)LBN;
#pragma unroll
        for (int q = 0; q < NSPLIT; ++q) {
          if (q == rank) continue;
          const uint32_t pa = mapa_u32(smem_u32(xch), (uint32_t)q) + (uint32_t)((par * NSPLIT + rank) * 128 + rowi) * 8u;
          asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(pa), "f"(mean), "f"(m2) : "memory");
          mbar_arrive_cluster(mapa_u32(smem_u32(&xbar[par]), (uint32_t)q));
        }
        mbar_wait_cluster(&xbar[par], (uint32_t)(tile_i >> 1) & 1u);
        float mq[NSPLIT], msum = 0.f, m2s = 0.f;
#pragma unroll
        for (int q = 0; q < NSPLIT; ++q) {
          const float2 o = q == rank ? make_float2(mean, m2) : xch[(par * NSPLIT + q) * 128 + rowi];
          mq[q] = o.x; msum += o.x; m2s += o.y;
        }
        mean = msum * (1.0f / NSPLIT);
        float dev = 0.f;
#pragma unroll
        for (int q = 0; q < NSPLIT; ++q) dev = fmaf(mq[q] - mean, mq[q] - mean, dev);
        var = (m2s + (float)LBN * dev) * (1.0f / 256.0f);
      } else if (NSPLIT > 1) {
        // (mean, M2) of this CTA's 128 columns → the peer's exchange slot; combine with the peer's (parallel-variance formula)
        const int slot = (tile_i & 1) * 128 + quad * 32 + lane;
        const float m2 = var * (float)LBN;
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(peer_xch + (uint32_t)slot * 8u), "f"(mean), "f"(m2) : "memory");
        mbar_arrive_cluster(peer_xbar);
        mbar_wait_cluster(xbar, xph);
        xph ^= 1;
        const float2 o = xch[slot];
        const float d = mean - o.x;
        var = (m2 + o.y + 0.5f * (float)LBN * d * d) * (1.0f / 256.0f);
        mean = 0.5f * (mean + o.x);
      }
      const float rstd = rsqrtf(var + p.eps);
      // ---------------- pass 2: normalise → slabs → TMA stores (units of 64 columns)
#pragma unroll 1
      for (int u = 0; u < LBN / 64; ++u) {
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int c = u * 2 + hf;
          tmem_ld32(taddr + c * 32, raw);
          tmem_ld_wait();
          const float* w = slnw + c * 32;
          const float* b = slnb + c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (__uint_as_float(raw[i]) - mean) * rstd * w[i] + b[i];
          if (p.has_c32) {
            uint8_t* rowp = (hf ? bufD : bufC) + lane * 128;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(rowp + ((i ^ swz) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          {
            uint8_t* rowp = bufA + lane * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<uint4*>(rowp + (((hf * 4 + i) ^ swz) << 4)) =
                  make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                             pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
          }
          if (p.has_c2) {
            uint8_t* rowp = bufB + lane * 128;
            const uint4* a4 = reinterpret_cast<const uint4*>(p.add2 + (size_t)((valid ? row : 0) % p.add2_period) * 256 + col0 + c * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 q = __ldg(a4 + i);
              const float2 a = unpack_bf16(q.x), bb = unpack_bf16(q.y), cc = unpack_bf16(q.z), d = unpack_bf16(q.w);
              *reinterpret_cast<uint4*>(rowp + (((hf * 4 + i) ^ swz) << 4)) =
                  make_uint4(pack_bf16(v[8 * i] + a.x, v[8 * i + 1] + a.y), pack_bf16(v[8 * i + 2] + bb.x, v[8 * i + 3] + bb.y),
                             pack_bf16(v[8 * i + 4] + cc.x, v[8 * i + 5] + cc.y), pack_bf16(v[8 * i + 6] + d.x, v[8 * i + 7] + d.y));
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tma_c, bufA, col0 + u * 64, row0);
          if (p.has_c32) { tma_store_2d(&tma_c32, bufC, col0 + u * 64, row0); tma_store_2d(&tma_c32, bufD, col0 + u * 64 + 32, row0); }
          if (p.has_c2) tma_store_2d(&tma_c2, bufB, col0 + u * 64, row0);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (NSPLIT > 1) cluster_sync_all();   // no CTA exits while its peer may still write into its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * LBN);
  }
}

CUtensorMap make_tmap_2d(const void* ptr, int rows, int cols, int ld, int box_rows, bool f32);
int device_sm_count();
void count_gemm_launch();

bool gemm_ln_supported(int N, int K, const GemmEpi& e) {
  if (e.ln_w == nullptr || N != 256 || K % 64 != 0) return false;
  if (e.mul || e.res || e.c_f32) return false;
  if (e.bias != nullptr && e.bias_period > 1) return false;
  return true;
}

template <int NSPLIT>
static void launch_ln(const CUtensorMap& ta, const bf16* W, int ldw, int K, const CUtensorMap& tres, const CUtensorMap& tc,
                      const CUtensorMap& tc32, const CUtensorMap& tc2, const LnParams& p, int num_m, cudaStream_t stream) {
  using Cfg = LnCfg<NSPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    VG_CUDA(cudaFuncSetAttribute(gemm_ln_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr_set = true;
  }
  CUtensorMap tb = make_tmap_2d(W, 256, K, ldw, Cfg::LBN, false);
  int groups = device_sm_count() / NSPLIT;
  if (groups > num_m) groups = num_m;
  launch_pdl(gemm_ln_kernel<NSPLIT>, dim3(groups * NSPLIT), dim3(192), Cfg::kSmem, stream, ta, tb, tres, tc, tc32, tc2, p);
}

void gemm_ln(const bf16* A, int lda, const bf16* W, int ldw, int M, int K, const GemmEpi& e, cudaStream_t stream) {
  VG_CHECK(gemm_ln_supported(256, K, e), "gemm_ln: unsupported epilogue");
  // (An eight-warp, register-resident LayerNorm epilogue — the one of ffn_fused.cu — was measured for the encoder's out-proj + LN1
  //  launch and dropped: 302.8 us vs 292.1 us for this kernel; at 5.1 TB/s the launch is bound by HBM, not by the epilogue.)
  LnParams p;
  p.bias = e.bias; p.ln_w = e.ln_w; p.ln_b = e.ln_b; p.add2 = e.add2; p.M = M; p.K = K; p.act = e.act;
  p.add2_period = e.add2_period > 0 ? e.add2_period : 1;
  p.has_res = e.res32 != nullptr; p.has_c32 = e.C32 != nullptr; p.has_c2 = e.C2 != nullptr; p.eps = e.ln_eps;
  VG_CHECK(!p.has_c2 || e.add2 != nullptr, "gemm_ln: C2 needs add2");
  CUtensorMap ta = make_tmap_2d(A, M, K, lda, LBM, false);
  CUtensorMap tc = make_tmap_2d(e.C, M, 256, e.ldc, 32, false);
  CUtensorMap tres = p.has_res ? make_tmap_2d(e.res32, M, 256, e.ldres32, 32, true) : tc;
  CUtensorMap tc32 = p.has_c32 ? make_tmap_2d(e.C32, M, 256, e.ldc32, 32, true) : tc;
  CUtensorMap tc2 = p.has_c2 ? make_tmap_2d(e.C2, M, 256, e.ldc2, 32, false) : tc;
  const int num_m = (M + LBM - 1) / LBM;
  // few row tiles and a long K (the decoders' 2048 -> 256 projections): split the columns over CTA pairs
  static int split_ok = -1;
  if (split_ok < 0) { const char* s = getenv("VGQA_LN_SPLIT"); split_ok = (s == nullptr || s[0] != '0') ? 1 : 0; }
  if (split_ok && num_m == 1 && K >= 256)
    // a single row tile (batch-1 decoders): a cluster of four CTAs streams a quarter of the K x 256 weights each
    launch_ln<4>(ta, W, ldw, K, tres, tc, tc32, tc2, p, num_m, stream);
  else if (split_ok && num_m * 2 <= device_sm_count() && K >= 512)
    launch_ln<2>(ta, W, ldw, K, tres, tc, tc32, tc2, p, num_m, stream);
  else
    launch_ln<1>(ta, W, ldw, K, tres, tc, tc32, tc2, p, num_m, stream);
  count_gemm_launch();
}

}  // namespace vg
