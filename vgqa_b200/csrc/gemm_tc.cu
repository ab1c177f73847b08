// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   C[M,N] = epilogue( A[M,K] · W[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// Replaces every nn.Linear / packed in-proj / out-proj / FFN matmul on the grounding hot path
// (reference: vgqa/core/decoder/modal_encoder.py:171-177, query_decoder.py:282-374,466-485,
// language/bert_module.py:59-141,196-225, model_utils.py:43-58).
//
// Structure (one CTA per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled A (128x64) and W (BNx64) tiles into a
//               STAGES-deep shared-memory ring, runs ahead across output tiles.
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128xBNx16), accumulators
//               double-buffered in TMEM (2 x BN fp32 columns) so the epilogue of tile i overlaps the
//               MMAs of tile i+1.
//   warps 2..5  epilogue: each thread owns one accumulator row (TMEM lane): tcgen05.ld 32 columns at a
//               time, bias / row-table, ReLU / erf-GELU, elementwise multiply, residual add and — when
//               BN == N == 256 — a full-row LayerNorm computed thread-locally (no shuffles), then
//               16-byte stores.
#include <mutex>
#include <unordered_map>

#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace vg {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom row

template <int BN>
struct GemmCfg {
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // power of two >= 32 for BN in {64,128,256}
  static constexpr int kStagingBytes = 4 * 2 * 4096;  // 4 epilogue warps x 2 buffers x (32 rows x 128 B)
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct EpiDev {
  void* C;
  bf16* C2;
  const bf16* add2;
  int ldc2, add2_period;
  float* C32;
  const float* res32;
  const float* bias;
  const bf16* mul;
  const bf16* res;
  const float* ln_w;
  const float* ln_b;
  int ldc, c_f32, bias_period, bias_ld, act, ldmul, ldres, ldres32, ldc32;
  float ln_eps;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// v[32] <- act(acc + bias) * mul + res   for row `row`, columns [col0, col0+32)
__device__ __forceinline__ void epi_transform(float (&v)[32], const EpiDev& ep, int row, int col0) {
  if (ep.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(ep.bias + (size_t)(row % ep.bias_period) * ep.bias_ld + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = __ldg(b4 + i);
      v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
  }
  if (ep.act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (ep.act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  }
  if (ep.mul != nullptr) {
    const uint4* m4 = reinterpret_cast<const uint4*>(ep.mul + (size_t)row * ep.ldmul + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u = __ldg(m4 + i);
      float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      v[8 * i + 0] *= a.x; v[8 * i + 1] *= a.y; v[8 * i + 2] *= b.x; v[8 * i + 3] *= b.y;
      v[8 * i + 4] *= c.x; v[8 * i + 5] *= c.y; v[8 * i + 6] *= d.x; v[8 * i + 7] *= d.y;
    }
  }
  if (ep.res != nullptr) {
    const uint4* r4 = reinterpret_cast<const uint4*>(ep.res + (size_t)row * ep.ldres + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u = __ldg(r4 + i);
      float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      v[8 * i + 0] += a.x; v[8 * i + 1] += a.y; v[8 * i + 2] += b.x; v[8 * i + 3] += b.y;
      v[8 * i + 4] += c.x; v[8 * i + 5] += c.y; v[8 * i + 6] += d.x; v[8 * i + 7] += d.y;
    }
  }
  if (ep.res32 != nullptr) {
    const float4* r4 = reinterpret_cast<const float4*>(ep.res32 + (size_t)row * ep.ldres32 + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 r = __ldg(r4 + i);
      v[4 * i + 0] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
    }
  }
}

__device__ __forceinline__ void epi_store(const float (&v)[32], const EpiDev& ep, int row, int col0) {
  if (ep.C32 != nullptr) {
    float4* c4 = reinterpret_cast<float4*>(ep.C32 + (size_t)row * ep.ldc32 + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) c4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  if (ep.c_f32) {
    float4* c4 = reinterpret_cast<float4*>(static_cast<float*>(ep.C) + (size_t)row * ep.ldc + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) c4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
    uint4* c4 = reinterpret_cast<uint4*>(static_cast<bf16*>(ep.C) + (size_t)row * ep.ldc + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      c4[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                         pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
  }
}

template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const EpiDev ep, int M, int N, int K) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint8_t* smem_stage_out = smem + Cfg::kStages * Cfg::kStageBytes;  // 1024-aligned (stage sizes are multiples of 1 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage_out + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM;
  const int num_n = N / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // PDL: the weight blocks of the first ring pass do not depend on the previous kernel — fetch them before waiting for it
      const int pre = (int)blockIdx.x < num_tiles ? (num_k < Cfg::kStages ? num_k : Cfg::kStages) : 0;
      for (int kb = 0; kb < pre; ++kb) {
        mbar_expect_tx(&full_bar[kb], Cfg::kStageBytes);
        tma_load_2d(smem_b + kb * Cfg::kBBytes, &tma_b, &full_bar[kb], kb * BK, ((int)blockIdx.x % num_n) * BN);
      }
      pdl_wait();
      pdl_launch_dependents();
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * BM;
        const int n0 = (tile % num_n) * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          if (first && kb < pre) {
            tma_load_2d(smem_a + stage * Cfg::kABytes, &tma_a, &full_bar[stage], kb * BK, m0);
          } else {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_2d(smem_a + stage * Cfg::kABytes, &tma_a, &full_bar[stage], kb * BK, m0);
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &tma_b, &full_bar[stage], kb * BK, n0);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);                       // frees the smem stage when the MMAs retire
          if (kb == num_k - 1) umma_commit(&tfull_bar[acc]);   // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    pdl_wait();   // mul / residual operands and the output buffer belong to earlier kernels of the stream
    const int quad = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool do_ln = ep.ln_w != nullptr;
    uint32_t nstore = 0;   // TMA stores this warp has issued: the slab alternates per STORE (a tile may have an odd number of units)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * BM;
      const int n0 = (tile % num_n) * BN;
      const int row = m0 + quad * 32 + lane;
      const bool valid = row < M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
      uint32_t raw[32];
      float v[32];
      if (!do_ln) {
        // TMEM → registers → 128B-swizzled smem slab (32 rows x 128 B per warp) → TMA store; double-buffered per warp
        uint8_t* my_stage = smem_stage_out + (warp - 2) * 8192;
        const int cols_per_unit = ep.c_f32 ? 32 : 64;   // 128 B of output per row
        const int units = BN / cols_per_unit;
#pragma unroll 1
        for (int u = 0; u < units; ++u) {
          uint8_t* buf = my_stage + (nstore & 1) * 4096;
          ++nstore;
          if (lane == 0) tma_store_wait_read<1>();      // the store issued two units ago has drained this buffer
          __syncwarp();
          const int halves = ep.c_f32 ? 1 : 2;
          for (int hf = 0; hf < halves; ++hf) {
            const int col = u * cols_per_unit + hf * 32;
            tmem_ld32(taddr + col, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
            if (valid) epi_transform(v, ep, row, n0 + col);
            uint8_t* rowp = buf + lane * 128;
            if (ep.c_f32) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(rowp + ((i ^ (lane & 7)) << 4)) =
                    make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                *reinterpret_cast<uint4*>(rowp + (((hf * 4 + i) ^ (lane & 7)) << 4)) =
                    make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                               pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tma_c, buf, n0 + u * cols_per_unit, m0 + quad * 32);
            tma_store_commit();
          }
        }
      } else {
        // pass 1: materialise pre-LN values back into TMEM, accumulate shifted moments
        float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          tmem_ld32(taddr + c * 32, raw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
          if (valid) epi_transform(v, ep, row, n0 + c * 32);
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
        const float inv_n = 1.0f / BN;
        const float dm = s1 * inv_n;
        const float mean = shift + dm;
        const float var = fmaxf(s2 * inv_n - dm * dm, 0.f);
        const float rstd = rsqrtf(var + ep.ln_eps);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          tmem_ld32(taddr + c * 32, raw);
          tmem_ld_wait();
          const float4* w4 = reinterpret_cast<const float4*>(ep.ln_w + n0 + c * 32);
          const float4* b4 = reinterpret_cast<const float4*>(ep.ln_b + n0 + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 w = __ldg(w4 + i), b = __ldg(b4 + i);
            v[4 * i + 0] = (__uint_as_float(raw[4 * i + 0]) - mean) * rstd * w.x + b.x;
            v[4 * i + 1] = (__uint_as_float(raw[4 * i + 1]) - mean) * rstd * w.y + b.y;
            v[4 * i + 2] = (__uint_as_float(raw[4 * i + 2]) - mean) * rstd * w.z + b.z;
            v[4 * i + 3] = (__uint_as_float(raw[4 * i + 3]) - mean) * rstd * w.w + b.w;
          }
          if (valid) {
            epi_store(v, ep, row, n0 + c * 32);
            if (ep.C2 != nullptr) {  // third output: LN(v) + positional rows (bf16)
              const uint4* a4 = reinterpret_cast<const uint4*>(ep.add2 + (size_t)(row % ep.add2_period) * 256 + n0 + c * 32);
              uint4* o4 = reinterpret_cast<uint4*>(ep.C2 + (size_t)row * ep.ldc2 + n0 + c * 32);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = __ldg(a4 + i);
                float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), cc = unpack_bf16(u.z), d = unpack_bf16(u.w);
                o4[i] = make_uint4(pack_bf16(v[8 * i] + a.x, v[8 * i + 1] + a.y), pack_bf16(v[8 * i + 2] + b.x, v[8 * i + 3] + b.y),
                                   pack_bf16(v[8 * i + 4] + cc.x, v[8 * i + 5] + cc.y), pack_bf16(v[8 * i + 6] + d.x, v[8 * i + 7] + d.y));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (warp >= 2 && lane == 0) tma_store_wait_all<0>();  // smem must outlive the bulk stores
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  VG_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  return fn;
}

void* tensor_map_encode_fn() { return reinterpret_cast<void*>(get_encode()); }

// 2D row-major [rows, cols] (bf16 or fp32) with row stride ld (elements); box = [box_rows x 128 bytes], 128B swizzle.
CUtensorMap make_tmap_2d(const void* ptr, int rows, int cols, int ld, int box_rows, bool f32) {
  CUtensorMap m;
  const size_t es = f32 ? 4 : 2;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  VG_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  VG_CHECK((ld * es) % 16 == 0, "TMA row stride must be a multiple of 16 bytes");
  CUresult r = get_encode()(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                            const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return m;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VGQA_PDL"); v = (e == nullptr || e[0] != '0') ? 1 : 0; }
  return v != 0;
}

static int g_num_sms = 0;
static int g_gemm_launches = 0;
int gemm_launch_count() { return g_gemm_launches; }
void count_gemm_launch() { ++g_gemm_launches; }
// Persistent kernels size their grids with device_sm_count().  The model caps it (sm_budget) for the encoder phase so that
// a few SMs stay free for the latency-bound decoder-phase launches of the previous batch running on the other stream.
static thread_local int g_sm_budget = 0;
void set_sm_budget(int n) { g_sm_budget = n; }
int device_sm_count() {
  if (g_num_sms == 0) {
    int dev = 0;
    VG_CUDA(cudaGetDevice(&dev));
    VG_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  return g_sm_budget > 0 ? g_sm_budget : g_num_sms;   // a budget above the SM count = short-lived CTAs (several waves)
}

template <int BN>
static void launch_gemm(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const EpiDev& ep,
                        cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    VG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  CUtensorMap ta = make_tmap_2d(A, M, K, lda, BM, false);
  CUtensorMap tb = make_tmap_2d(W, N, K, ldw, BN, false);
  CUtensorMap tc = make_tmap_2d(ep.C, M, N, ep.ldc, 32, ep.c_f32 != 0);
  const int tiles = ((M + BM - 1) / BM) * (N / BN);
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  launch_pdl(gemm_tc_kernel<BN>, dim3(grid), dim3(192), Cfg::kSmemBytes, stream, ta, tb, tc, ep, M, N, K);
  VG_CUDA(cudaGetLastError());
  ++g_gemm_launches;
}

void gemm_bf16_tn(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const GemmEpi& e,
                  cudaStream_t stream) {
  VG_CHECK(M > 0 && N > 0 && K > 0, "gemm: empty problem");
  VG_CHECK(K % BK == 0, "gemm: K must be a multiple of 64");
  VG_CHECK(N % 64 == 0, "gemm: N must be a multiple of 64");
  VG_CHECK(e.C != nullptr && e.ldc % 8 == 0, "gemm: bad output");
  if (g_num_sms == 0) {
    int dev = 0;
    VG_CUDA(cudaGetDevice(&dev));
    VG_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (gemm_ln_supported(N, K, e)) {
    gemm_ln(A, lda, W, ldw, M, K, e, stream);
    return;
  }
  if (M > BM && gemm_ws_supported(N, K, e)) {   // a single row tile: 64-column streaming tiles below put more SMs on the weights
    gemm_ws(A, nullptr, 0, lda, W, ldw, M, N, K, e, stream);
    return;
  }
  EpiDev ep;
  ep.C2 = e.C2; ep.ldc2 = e.ldc2; ep.add2 = e.add2; ep.add2_period = e.add2_period > 0 ? e.add2_period : 1;
  ep.C = e.C; ep.bias = e.bias; ep.mul = e.mul; ep.res = e.res; ep.ln_w = e.ln_w; ep.ln_b = e.ln_b;
  ep.ldc = e.ldc; ep.c_f32 = e.c_f32; ep.bias_period = e.bias_period > 0 ? e.bias_period : 1;
  ep.bias_ld = e.bias_ld; ep.act = e.act; ep.ldmul = e.ldmul; ep.ldres = e.ldres; ep.ln_eps = e.ln_eps;
  ep.C32 = e.C32; ep.ldc32 = e.ldc32; ep.res32 = e.res32; ep.ldres32 = e.ldres32;
  if (e.ln_w != nullptr) {
    VG_CHECK(N == 256, "gemm: the fused LayerNorm epilogue needs N == 256");
    launch_gemm<256>(A, lda, W, ldw, M, N, K, ep, stream);
  } else if (M <= BM && N % 64 == 0) {
    // a single row tile (the text tower's handful of tokens, batch-1 heads): the launch is bound by how fast ONE SM can stream its
    // K x BN weight slice — 64-column tiles put four times as many SMs on the weights as 256-column ones
    launch_gemm<64>(A, lda, W, ldw, M, N, K, ep, stream);
  } else if (N % 256 == 0) {
    launch_gemm<256>(A, lda, W, ldw, M, N, K, ep, stream);
  } else if (N % 128 == 0) {
    launch_gemm<128>(A, lda, W, ldw, M, N, K, ep, stream);
  } else {
    launch_gemm<64>(A, lda, W, ldw, M, N, K, ep, stream);
  }
}

}  // namespace vg
