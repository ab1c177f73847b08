// Weight-stationary tcgen05 GEMM for the K <= 256 (BN = 256) / K <= 512 (BN = 128) Linear layers of the encoder
// (fused Q|K|V in-projection, FFN linear1) and of the decoders' query side.
//
//   C[M,N] = act( A[M,K] · W[N,K]^T + bias[N] ),  bf16 operands, fp32 accumulation in TMEM, bf16 or fp32 output.
//
// Why a second GEMM kernel: with K = 256 a 128 x 256 output tile needs only 16 UMMA instructions (2048 tensor
// cycles) but 192 KB of operands; re-streaming the 128 KB weight slice for every tile made L2→SM traffic and the
// epilogue the limiters (profiles/r01_*).  Here each CTA owns ONE n-tile for its whole life:
//   * its W slice (BN x K bf16, <= 128 KB) is TMA-loaded once and stays resident in shared memory,
//   * only the 128 x 64 A blocks stream through a 3-stage ring (16 KB each),
//   * the bias slice sits in shared memory,
//   * 8 epilogue warps (two per TMEM lane quadrant, each taking half of the columns) drain the double-buffered
//     accumulator: tcgen05.ld → bias/activation → bf16 pack → 128B-swizzled smem slab → TMA store.
// Optional second A operand: n-tiles < a_switch read A, the others read A2 (same shape) — used for the encoder's
// in-projection where Q,K are computed from (x + pos) and V from x (modal_encoder.py:171-172).
#include "common.h"
#include "ptx.cuh"

namespace vg {

static constexpr int WBM = 128;
static constexpr int WBK = 64;
static constexpr int kWsStages = 4;
static constexpr int kWsEpiWarps = 8;

struct WsParams {
  const float* bias;  // [N] or nullptr
  int M, N, K;
  int act;
  int a_switch;  // n-tiles >= a_switch read the second A operand
};

template <int BN>
struct WsCfg {
  static constexpr int kWBytesMax = 128 * 1024;
  static constexpr int kABytes = WBM * WBK * 2;
  static constexpr int kStaging = kWsEpiWarps * 4096;
  static constexpr int kSmem = kWBytesMax + kWsStages * kABytes + kStaging + BN * 4 + 1024 + 256;
};

__device__ __forceinline__ float ws_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN>
__global__ void __launch_bounds__(320, 1)
gemm_ws_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_a2,
               const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_c, const WsParams p) {
  using Cfg = WsCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;                                     // [K/64][BN x 128 B]
  uint8_t* smem_a = smem + Cfg::kWBytesMax;                   // ring of 128 x 128 B blocks
  uint8_t* smem_out = smem_a + kWsStages * Cfg::kABytes;      // 8 x 4 KB slabs
  float* sbias = reinterpret_cast<float*>(smem_out + Cfg::kStaging);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sbias) + BN * 4);
  uint64_t* full_bar = bars;                 // [stages]
  uint64_t* empty_bar = bars + kWsStages;    // [stages]
  uint64_t* tfull_bar = empty_bar + kWsStages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint64_t* w_bar = tempty_bar + 2;             // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.M + WBM - 1) / WBM;
  const int num_n = p.N / BN;
  const int num_k = p.K / WBK;
  const int n_tile = blockIdx.x % num_n;
  const int cta_in_n = blockIdx.x / num_n;
  const int ctas_per_n = gridDim.x / num_n;
  const int n0 = n_tile * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_a2); tma_prefetch_desc(&tma_w); tma_prefetch_desc(&tma_c);
    for (int s = 0; s < kWsStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kWsEpiWarps); }
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  if (warp >= 2) {  // bias slice → smem (zeros when there is no bias)
    for (int i = threadIdx.x - 64; i < BN; i += kWsEpiWarps * 32) sbias[i] = p.bias ? p.bias[n0 + i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_bar, (uint32_t)(BN * p.K * 2));
      for (int kb = 0; kb < num_k; ++kb) tma_load_2d(smem_w + kb * (BN * 128), &tma_w, w_bar, kb * WBK, n0);
      pdl_wait();                // PDL: the resident weight slice was fetched while the previous kernel was still running
      pdl_launch_dependents();
      const CUtensorMap* ta = n_tile < p.a_switch ? &tma_a : &tma_a2;
      int stage = 0;
      uint32_t phase = 0;
      for (int mt = cta_in_n; mt < num_m; mt += ctas_per_n) {
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kABytes);
          tma_load_2d(smem_a + stage * Cfg::kABytes, ta, &full_bar[stage], kb * WBK, mt * WBM);
          if (++stage == kWsStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(WBM, BN);
    mbar_wait(w_bar, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int mt = cta_in_n; mt < num_m; mt += ctas_per_n) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smem_w + kb * (BN * 128)));
#pragma unroll
          for (int k = 0; k < WBK / 16; ++k)
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == num_k - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == kWsStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue: warps 2..9 =====================
    pdl_wait();   // the output buffer may still be read by the previous kernel of the stream
    const int quad = warp & 3;              // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;       // which half of the BN columns
    uint8_t* slab = smem_out + (warp - 2) * 4096;
    const int col_base = half * (BN / 2);
    constexpr int kUnits = (BN / 2) / 64;   // 64 bf16 columns = one 128-byte slab row
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int mt = cta_in_n; mt < num_m; mt += ctas_per_n) {
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + col_base;
      uint32_t raw[32];
#pragma unroll
      for (int u = 0; u < kUnits; ++u) {
        if (lane == 0) tma_store_wait_read<0>();  // this warp's slab is free again
        __syncwarp();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int c = u * 2 + hf;
          tmem_ld32(taddr + c * 32, raw);   // LDTM latency is ~12 cycles; the second warp of the SMSP covers it
          tmem_ld_wait();
          float v[32];
          const float* sb = sbias + col_base + c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) + sb[i];
          if (p.act == ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          } else if (p.act == ACT_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = ws_gelu(v[i]);
          }
          uint8_t* rowp = slab + lane * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(rowp + (((hf * 4 + i) ^ (lane & 7)) << 4)) =
                make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                           pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tma_c, slab, n0 + col_base + u * 64, mt * WBM + quad * 32);
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------------ host
CUtensorMap make_tmap_2d(const void* ptr, int rows, int cols, int ld, int box_rows, bool f32);  // gemm_tc.cu
int device_sm_count();
void count_gemm_launch();

template <int BN>
static void launch_ws(const bf16* A, const bf16* A2, int lda, const bf16* W, int ldw, int M, int N, int K,
                      const WsParams& p, void* C, int ldc, cudaStream_t stream) {
  using Cfg = WsCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    VG_CUDA(cudaFuncSetAttribute(gemm_ws_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr_set = true;
  }
  CUtensorMap ta = make_tmap_2d(A, M, K, lda, WBM, false);
  CUtensorMap ta2 = make_tmap_2d(A2 ? A2 : A, M, K, lda, WBM, false);
  CUtensorMap tw = make_tmap_2d(W, N, K, ldw, BN, false);
  CUtensorMap tc = make_tmap_2d(C, M, N, ldc, 32, false);
  const int num_n = N / BN, num_m = (M + WBM - 1) / WBM;
  int per_n = device_sm_count() / num_n;
  if (per_n < 1) per_n = 1;
  if (per_n > num_m) per_n = num_m;
  launch_pdl(gemm_ws_kernel<BN>, dim3(num_n * per_n), dim3(320), Cfg::kSmem, stream, ta, ta2, tw, tc, p);
  VG_CUDA(cudaGetLastError());
  count_gemm_launch();
}

bool gemm_ws_supported(int N, int K, const GemmEpi& e) {
  if (e.ln_w || e.mul || e.res || e.res32 || e.C32 || e.C2 || e.c_f32) return false;
  if (e.bias != nullptr && e.bias_period > 1) return false;
  if (K % 64 != 0) return false;
  if (N % 256 == 0 && K <= 256) return true;
  if (N % 128 == 0 && K <= 512) return true;
  return false;
}

// C = act(A W^T + bias); n-tiles (of 256 or 128 columns) with index >= a_switch read A2 instead of A.
void gemm_ws(const bf16* A, const bf16* A2, int a_switch_col, int lda, const bf16* W, int ldw, int M, int N, int K,
             const GemmEpi& e, cudaStream_t stream) {
  VG_CHECK(gemm_ws_supported(N, K, e), "gemm_ws: unsupported problem");
  WsParams p;
  p.bias = e.bias; p.M = M; p.N = N; p.K = K; p.act = e.act;
  if (N % 256 == 0 && K <= 256) {
    p.a_switch = A2 ? a_switch_col / 256 : (N / 256);
    launch_ws<256>(A, A2, lda, W, ldw, M, N, K, p, e.C, e.ldc, stream);
  } else {
    p.a_switch = A2 ? a_switch_col / 128 : (N / 128);
    launch_ws<128>(A, A2, lda, W, ldw, M, N, K, p, e.C, e.ldc, stream);
  }
}

}  // namespace vg
