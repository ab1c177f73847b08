// The ResNet101 extractor on this library's kernels: `self.vis_encoder[0].body` of VSTGNet = torchvision
// resnet101 with FrozenBatchNorm2d, layer4 output (vgqa/core/vision/backbone.py:13-57,60-113; grounding_net.py:49,99).
//
// Layout.  Activations are channels-last bf16 in a SPATIALLY PADDED frame grid [n, H + 2, W + 2, C] whose border rows are zero.
// In that layout a 3x3 / stride 1 / pad 1 convolution is a GEMM whose A operand of tap (ky, kx) is the SAME matrix shifted by
// (ky - 1) * (W + 2) + (kx - 1) rows: the TMA producer adds that row offset per k-block (rows before the first / after the last are
// zero-filled by the tensor map), so there is no im2col buffer — K = 9 C runs over (tap, channel block) and the accumulator stays in
// TMEM.  Outputs at border positions are computed from neighbours and discarded: the epilogue writes zeros there, which keeps the
// invariant for the next 3x3.  The 1x1 convolutions are plain GEMMs over the same rows.  The price is (H+2)(W+2)/(HW) more rows
// (7 % at 56x56, 31 % at 14x14); what it buys is 9x less A traffic than an explicit im2col.  FrozenBN is folded into the weights and
// the bias at pack time; ReLU, the residual add and the border mask are the GEMM epilogue.
//
// Kernels: conv_gemm (tcgen05 / TMEM / TMA, 128 x {64,128,256} tiles, persistent, double-buffered accumulator), the stem's im2col
// (7x7 / 2 on 3 input channels: one CTA per output row, input rows staged in shared memory), 3x3 / 2 max-pool, and — for the three
// stride-2 blocks only — an explicit 3x3 im2col and a 2x row subsample into the next resolution's padded grid.
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"
#include "resnet.h"

namespace vg {

CUtensorMap make_tmap_2d(const void* ptr, int rows, int cols, int ld, int box_rows, bool f32);  // gemm_tc.cu
int device_sm_count();

// ------------------------------------------------------------------------------------------------ the convolution GEMM
struct ConvParams {
  const float* bias;   // [N]
  const bf16* res;     // [M, ldres] residual added before the ReLU, or nullptr
  int ldres;
  int M, N, K;         // K = taps * C
  int relu;
  int taps, kpt, Wp;   // kpt = 64-wide k-blocks per tap; tap t reads A rows shifted by (t / 3 - 1) * Wp + (t % 3 - 1) when taps == 9
  int Hp;              // rows per frame = Hp * Wp, border rows are written as zeros; 0 = plain row-major output
};

// DEEP: one more pipeline stage instead of the second staging slab per epilogue warp — for long-K GEMMs without a residual (the 3x3
// convolutions, the 1x1 reductions of layer3 / layer4), whose time is the main loop: three 48 KB stages hold 0.8 us of MMA work, less
// than the latency of a TMA load
template <int BN, bool DEEP>
struct ConvCfg {
  static constexpr int kStages = BN == 256 ? (DEEP ? 4 : 3) : (BN == 128 ? (DEEP ? 6 : 5) : 8);
  static constexpr int kABytes = 128 * 64 * 2;
  static constexpr int kBBytes = BN * 64 * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiWarps = BN >= 128 ? 8 : 4;   // two warps per TMEM lane quadrant (half of the tile's columns each) from 128 columns on
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kSlabs = DEEP ? 1 : 2;             // staging slabs of 32 rows x 128 B per epilogue warp
  static constexpr int kStaging = kEpiWarps * kSlabs * 4096;
  static constexpr int kSmem = kStages * kStageBytes + kStaging + 1024 + 512;   // + alignment slack + barriers
  static_assert(kSmem <= 232448, "conv_gemm smem");
};

template <int BN, bool DEEP>
__global__ void __launch_bounds__(ConvCfg<BN, DEEP>::kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_r, const ConvParams p) {
  using Cfg = ConvCfg<BN, DEEP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint8_t* smem_out = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + Cfg::kStaging);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;              // [epilogue warps][2]: the residual segment of a store unit has landed in the slab
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * Cfg::kEpiWarps);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.M + 127) / 128, num_n = p.N / BN, num_tiles = num_m * num_n, num_k = p.K / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_b); tma_prefetch_desc(&tma_c); tma_prefetch_desc(&tma_r);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2 * Cfg::kEpiWarps; ++s) mbar_init(&res_bar[s], 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], Cfg::kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      pdl_wait();
      pdl_launch_dependents();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * 128, n0 = (tile % num_n) * BN;
        int tap = 0, kc = 0;
        for (int kb = 0; kb < num_k; ++kb) {
          const int off = p.taps == 9 ? (tap / 3 - 1) * p.Wp + (tap % 3 - 1) : 0;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_a + stage * Cfg::kABytes, &tma_a, &full_bar[stage], kc * 64, m0 + off);   // rows < 0 or >= M: zeros
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tma_b, &full_bar[stage], kb * 64, n0);
          if (++kc == p.kpt) { kc = 0; ++tap; }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
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
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == num_k - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue: bias, residual, ReLU, border mask → bf16 slab → TMA store =====================
    // warps 2..5 (and 6..9 from 128 columns on): a warp owns the TMEM lane quadrant warp % 4 and one half of the tile's columns
    pdl_wait();
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;                          // which half of the columns (always 0 with four warps)
    constexpr int kUnits = BN / 64 / (Cfg::kEpiWarps / 4);     // 64-column store units per warp and tile
    uint8_t* my_stage = smem_out + (warp - 2) * (Cfg::kSlabs * 4096);
    uint64_t* my_rbar = res_bar + 2 * (warp - 2);
    int acc = 0;
    uint32_t acc_phase = 0, q = 0;   // q: store units this warp has processed (slab = q & 1)
    const int fr = p.Hp * p.Wp;
    const bool has_res = p.res != nullptr;
    // The residual segment of a unit (32 rows x 128 bytes, the same box as the store) is brought INTO the unit's slab by TMA, one unit
    // ahead; the threads add their row in place and the slab goes out again.  unit q of this warp = (tile q / kUnits, unit q % kUnits)
    auto unit_coords = [&](uint32_t qq, int& col0, int& row0) -> bool {
      const int tile = (int)blockIdx.x + (int)(qq / kUnits) * (int)gridDim.x;
      if (tile >= num_tiles) return false;
      col0 = (tile % num_n) * BN + part * kUnits * 64 + (int)(qq % kUnits) * 64;
      row0 = (tile / num_n) * 128 + quad * 32;
      return true;
    };
    auto prefetch_res = [&](uint32_t qq) {   // lane 0
      int col0, row0;
      if (!unit_coords(qq, col0, row0)) return;
      mbar_expect_tx(&my_rbar[qq & 1], 4096);
      tma_load_2d(my_stage + (qq & 1) * 4096, &tma_r, &my_rbar[qq & 1], col0, row0);
    };
    if (has_res && lane == 0) prefetch_res(0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * 128, n0 = (tile % num_n) * BN + part * kUnits * 64;
      const int row = m0 + quad * 32 + lane;
      bool live = row < p.M;
      if (fr > 0 && live) {
        const int rr = row % fr, hp = rr / p.Wp, wp = rr - hp * p.Wp;
        live = hp != 0 && hp != p.Hp - 1 && wp != 0 && wp != p.Wp - 1;
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + part * kUnits * 64;
#pragma unroll 1
      for (int u = 0; u < kUnits; ++u, ++q) {
        uint8_t* buf = my_stage + (Cfg::kSlabs == 2 ? (q & 1) * 4096 : 0);
        if (lane == 0) {
          if (has_res) {
            tma_store_wait_read<0>();      // the other slab's store (unit q - 1) has been read out: the next residual may land in it
            prefetch_res(q + 1);
          } else {
            tma_store_wait_read<Cfg::kSlabs - 1>();      // the previous store out of this slab has drained it
          }
        }
        __syncwarp();
        if (has_res) mbar_wait(&my_rbar[q & 1], (q >> 1) & 1);
        uint8_t* rowp = buf + lane * 128;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int col = n0 + u * 64 + hf * 32;
          uint32_t raw[32];
          tmem_ld32(taddr + u * 64 + hf * 32, raw);
          tmem_ld_wait();
          float v[32];
          if (live) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(b4 + i);
              v[4 * i] = __uint_as_float(raw[4 * i]) + b.x; v[4 * i + 1] = __uint_as_float(raw[4 * i + 1]) + b.y;
              v[4 * i + 2] = __uint_as_float(raw[4 * i + 2]) + b.z; v[4 * i + 3] = __uint_as_float(raw[4 * i + 3]) + b.w;
            }
            if (has_res) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 rq = *reinterpret_cast<const uint4*>(rowp + (((hf * 4 + i) ^ (lane & 7)) << 4));
                const float2 a = unpack_bf16(rq.x), b = unpack_bf16(rq.y), c = unpack_bf16(rq.z), d = unpack_bf16(rq.w);
                v[8 * i] += a.x; v[8 * i + 1] += a.y; v[8 * i + 2] += b.x; v[8 * i + 3] += b.y;
                v[8 * i + 4] += c.x; v[8 * i + 5] += c.y; v[8 * i + 6] += d.x; v[8 * i + 7] += d.y;
              }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(rowp + (((hf * 4 + i) ^ (lane & 7)) << 4)) =
                make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                           pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tma_c, buf, n0 + u * 64, m0 + quad * 32);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all<0>();   // shared memory must outlive the bulk stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN, bool DEEP>
static void launch_conv(const bf16* A, int C, const bf16* W, const ConvParams& p, bf16* out, cudaStream_t st) {
  using Cfg = ConvCfg<BN, DEEP>;
  static bool attr_set = false;
  if (!attr_set) {
    VG_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr_set = true;
  }
  CUtensorMap ta = make_tmap_2d(A, p.M, C, C, 128, false);
  CUtensorMap tb = make_tmap_2d(W, p.N, p.K, p.K, BN, false);
  CUtensorMap tc = make_tmap_2d(out, p.M, p.N, p.N, 32, false);
  CUtensorMap tr = make_tmap_2d(p.res != nullptr ? (const void*)p.res : (const void*)out, p.M, p.N, p.ldres > 0 ? p.ldres : p.N, 32, false);
  const int tiles = ((p.M + 127) / 128) * (p.N / BN);
  const int grid = std::min(tiles, device_sm_count());
  launch_pdl(conv_gemm_kernel<BN, DEEP>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmem, st, ta, tb, tc, tr, p);
  VG_CUDA(cudaGetLastError());
}

// out[M, O] = epilogue( sum_taps A[rows shifted by the tap, C] · W[O, tap * C + c]^T )
static void conv_gemm(const bf16* A, const ResNet::Conv& cv, int taps, int M, int Hp, int Wp, const bf16* res, bool relu, bf16* out,
                      cudaStream_t st) {
  const int C = cv.C * cv.taps / taps;   // an explicit im2col operand carries its taps inside C
  VG_CHECK(C % 64 == 0 && cv.O % 64 == 0, "conv_gemm: channel counts must be multiples of 64");
  ConvParams p;
  p.bias = cv.bias; p.res = res; p.ldres = cv.O; p.M = M; p.N = cv.O; p.K = cv.C * cv.taps; p.relu = relu ? 1 : 0;
  p.taps = taps; p.kpt = C / 64; p.Wp = Wp; p.Hp = Hp;
  const bool deep = res == nullptr && p.K >= 1024;
  if (cv.O % 256 == 0) { if (deep) launch_conv<256, true>(A, C, cv.W, p, out, st); else launch_conv<256, false>(A, C, cv.W, p, out, st); }
  else if (cv.O % 128 == 0) { if (deep) launch_conv<128, true>(A, C, cv.W, p, out, st); else launch_conv<128, false>(A, C, cv.W, p, out, st); }
  else launch_conv<64, false>(A, C, cv.W, p, out, st);
}

// ------------------------------------------------------------------------------------------------ the other kernels
// Stem operand: frames NCHW fp32 [n, 3, R, R] → A0 [n * Ho * Wo, 192] bf16 (Ho = Wo = R / 2), column c * 49 + ky * 7 + kx =
// frame[c, 2 oy - 3 + ky, 2 ox - 3 + kx] (zero outside), columns 147..191 zero.  One CTA per output row: the 3 x 7 input rows it
// needs are staged (zero-padded by 3 on both sides) in shared memory.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ fr, bf16* __restrict__ A, int R) {
  extern __shared__ float srow[];   // [3][7][R + 6]
  __shared__ int koff[192];         // operand column → offset of its tap in srow (relative to 2 ox), -1 = padding column
  const int Ho = R >> 1, Rp = R + 6;
  const int n = blockIdx.x / Ho, oy = blockIdx.x - n * Ho;
  if (threadIdx.x < 192) {
    const int k = threadIdx.x, c = k / 49, r = k - c * 49, ky = r / 7, kx = r - ky * 7;
    koff[k] = k < 147 ? (c * 7 + ky) * Rp + kx : -1;
  }
  for (int cr = threadIdx.x >> 5; cr < 21; cr += 8) {   // one warp per (channel, ky) row: coalesced reads
    const int c = cr / 7, ky = cr - c * 7, iy = 2 * oy - 3 + ky;
    const float* src = fr + (((size_t)n * 3 + c) * R + iy) * R;
    const bool in = iy >= 0 && iy < R;
    for (int xx = threadIdx.x & 31; xx < Rp; xx += 32) {
      const int ix = xx - 3;
      srow[cr * Rp + xx] = (in && ix >= 0 && ix < R) ? __ldg(src + ix) : 0.f;
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(A + ((size_t)n * Ho + oy) * Ho * 192);
  for (int i = threadIdx.x; i < Ho * 24; i += blockDim.x) {
    const int ox = i / 24, j = i - ox * 24;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int o = koff[j * 8 + e];
      v[e] = o >= 0 ? srow[o + 2 * ox] : 0.f;
    }
    dst[i] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

__device__ __forceinline__ uint4 max_bf16x8(uint4 a, uint4 b) {
  const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
  const float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
  return make_uint4(pack_bf16(fmaxf(a0.x, b0.x), fmaxf(a0.y, b0.y)), pack_bf16(fmaxf(a1.x, b1.x), fmaxf(a1.y, b1.y)),
                    pack_bf16(fmaxf(a2.x, b2.x), fmaxf(a2.y, b2.y)), pack_bf16(fmaxf(a3.x, b3.x), fmaxf(a3.y, b3.y)));
}
// MaxPool2d(3, stride 2, padding 1): in [n, Hi, Hi, C] → out padded [n, Hi/2 + 2, Hi/2 + 2, C] (zero border); 16-byte chunks
__global__ void __launch_bounds__(256) maxpool_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int Hi, int cpr, long long total) {
  const int Hq = Hi >> 1, Hp = Hq + 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cpr;
    const int j = (int)(i - r * cpr);
    const long long n = r / (Hp * Hp);
    const int rr = (int)(r - n * Hp * Hp), hp = rr / Hp, wp = rr - hp * Hp;
    uint4 m = make_uint4(0u, 0u, 0u, 0u);
    if (hp != 0 && hp != Hp - 1 && wp != 0 && wp != Hp - 1) {
      const int h0 = 2 * (hp - 1) - 1, w0 = 2 * (wp - 1) - 1;
      bool first = true;
      for (int dy = 0; dy < 3; ++dy) {
        const int h = h0 + dy;
        if (h < 0 || h >= Hi) continue;
        for (int dx = 0; dx < 3; ++dx) {
          const int w = w0 + dx;
          if (w < 0 || w >= Hi) continue;
          const uint4 q = __ldg(in + ((n * Hi + h) * Hi + w) * cpr + j);
          m = first ? q : max_bf16x8(m, q);
          first = false;
        }
      }
    }
    out[i] = m;
  }
}
// 3x3 / stride 2 / pad 1 operand: in padded [n, Hi + 2, Hi + 2, C] → A [n * (Hi/2 + 2)^2, 9 C] (rows of the output's padded grid,
// zero on its border), column tap * C + c.  One 16-byte chunk per thread.
__global__ void __launch_bounds__(256) im2col3x3_s2_kernel(const uint4* __restrict__ in, uint4* __restrict__ A, int Hi, int cpr, long long total) {
  const int Hpi = Hi + 2, Hq = Hi >> 1, Hp = Hq + 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (9 * cpr);
    const int t = (int)(i - r * 9 * cpr), tap = t / cpr, j = t - tap * cpr;
    const long long n = r / (Hp * Hp);
    const int rr = (int)(r - n * Hp * Hp), hp = rr / Hp, wp = rr - hp * Hp;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (hp != 0 && hp != Hp - 1 && wp != 0 && wp != Hp - 1) {
      const int hi = 2 * (hp - 1) + tap / 3, wi = 2 * (wp - 1) + tap % 3;   // padded input coordinates of (2 ho - 1 + ky, 2 wo - 1 + kx)
      v = __ldg(in + ((n * Hpi + hi) * Hpi + wi) * cpr + j);
    }
    A[i] = v;
  }
}
// rows of every second position: in padded [n, Hi + 2, Hi + 2, C] → out padded [n, Hi/2 + 2, Hi/2 + 2, C] (the 1x1 / stride 2 downsample)
__global__ void __launch_bounds__(256) subsample_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int Hi, int cpr, long long total) {
  const int Hpi = Hi + 2, Hq = Hi >> 1, Hp = Hq + 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cpr;
    const int j = (int)(i - r * cpr);
    const long long n = r / (Hp * Hp);
    const int rr = (int)(r - n * Hp * Hp), hp = rr / Hp, wp = rr - hp * Hp;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (hp != 0 && hp != Hp - 1 && wp != 0 && wp != Hp - 1)
      v = __ldg(in + ((n * Hpi + 2 * (hp - 1) + 1) * Hpi + 2 * (wp - 1) + 1) * cpr + j);
    out[i] = v;
  }
}
// interior of a padded map: [n, H + 2, H + 2, C] bf16 → [n, H, H, C] bf16 and / or fp32
__global__ void __launch_bounds__(256) unpad_kernel(const uint4* __restrict__ in, uint4* __restrict__ o16, float4* __restrict__ o32, int H, int cpr,
                                                    long long total) {
  const int Hp = H + 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cpr;
    const int j = (int)(i - r * cpr);
    const long long n = r / (H * H);
    const int rr = (int)(r - n * H * H), h = rr / H, w = rr - h * H;
    const uint4 q = __ldg(in + ((n * Hp + h + 1) * Hp + w + 1) * cpr + j);
    if (o16 != nullptr) o16[i] = q;
    if (o32 != nullptr) {
      const float2 a = unpack_bf16(q.x), b = unpack_bf16(q.y), c = unpack_bf16(q.z), d = unpack_bf16(q.w);
      o32[2 * i] = make_float4(a.x, a.y, b.x, b.y);
      o32[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
    }
  }
}

static int grid_for(long long n) { return (int)std::min<long long>((n + 255) / 256, 148 * 32); }

// ------------------------------------------------------------------------------------------------ weights
void ResNet::pack(const HasFn& has, const GetFn& get, const std::function<bf16*(const float*, size_t)>& to_bf16,
                  const std::function<float*(const float*, size_t)>& to_f32) {
  const std::string root = "vis_encoder.0.body.";
  if (!has(root + "conv1.weight")) return;
  // conv [O, C, k, k] + FrozenBN (backbone.py:47-57: scale = w * rsqrt(var + 1e-5), bias = b - mean * scale) → W [O, kpad] + bias [O]
  auto fold = [&](const std::string& conv, const std::string& bn, int O, int C, int k, Conv& out) {
    const float* w = get(conv + ".weight", {O, C, k, k});
    const float* g = get(bn + ".weight", {O});
    const float* b = get(bn + ".bias", {O});
    const float* rm = get(bn + ".running_mean", {O});
    const float* rv = get(bn + ".running_var", {O});
    const int taps = k * k;
    const bool is_stem = k == 7;
    const int kp = is_stem ? 192 : taps * C;
    std::vector<float> W((size_t)O * kp, 0.f), bias(O);
    for (int o = 0; o < O; ++o) {
      const float scale = g[o] / std::sqrt(rv[o] + 1e-5f);
      bias[o] = b[o] - rm[o] * scale;
      for (int c = 0; c < C; ++c)
        for (int t = 0; t < taps; ++t) {
          const float v = w[((size_t)o * C + c) * taps + t] * scale;
          if (is_stem) W[(size_t)o * kp + c * taps + t] = v;   // the stem's operand is ordered (channel, ky, kx)
          else W[(size_t)o * kp + (size_t)t * C + c] = v;       // the others (tap, channel)
        }
    }
    out.W = to_bf16(W.data(), W.size());
    out.bias = to_f32(bias.data(), bias.size());
    out.O = O; out.C = is_stem ? 192 : C; out.taps = is_stem ? 1 : taps;
  };
  fold(root + "conv1", root + "bn1", 64, 3, 7, stem);
  int cin = 64;
  for (int l = 0; l < kLayers; ++l) {
    const int width = 64 << l;
    layer[l].resize(blocks[l]);
    for (int b = 0; b < blocks[l]; ++b) {
      const std::string p = root + "layer" + std::to_string(l + 1) + "." + std::to_string(b) + ".";
      Block& k = layer[l][b];
      k.stride = (b == 0 && l > 0) ? 2 : 1;
      fold(p + "conv1", p + "bn1", width, cin, 1, k.c1);
      fold(p + "conv2", p + "bn2", width, width, 3, k.c2);
      fold(p + "conv3", p + "bn3", 4 * width, width, 1, k.c3);
      k.has_down = b == 0;
      if (k.has_down) fold(p + "downsample.0", p + "downsample.1", 4 * width, cin, 1, k.down);
      cin = 4 * width;
    }
  }
  loaded = true;
}

// ------------------------------------------------------------------------------------------------ workspace
void ResNet::ensure_workspace(int frames, int R) {
  if (frames <= cap_frames && R <= cap_R && x != nullptr) return;
  frames = std::max(frames, cap_frames); R = std::max(R, cap_R);
  release();
  const size_t nf = (size_t)frames, Ho = R / 2;
  size_t mx = 0, mo1 = 0, mo2 = 0, ma2 = 0, mxs = 0;
  int H = R / 4, cin = 64;
  for (int l = 0; l < kLayers; ++l) {
    const size_t width = (size_t)64 << l;
    if (l > 0) {
      const size_t rows_in = nf * (H + 2) * (H + 2);
      H /= 2;
      const size_t rows = nf * (H + 2) * (H + 2);
      mo1 = std::max(mo1, rows_in * width);
      ma2 = std::max(ma2, rows * 9 * width);
      mxs = std::max(mxs, rows * cin);
    }
    const size_t rows = nf * (H + 2) * (H + 2);
    mx = std::max(mx, rows * 4 * width);
    mo1 = std::max(mo1, rows * width);
    mo2 = std::max(mo2, rows * width);
    cin = 4 * (int)width;
  }
  const size_t pad = 4096;
  VG_CUDA(cudaMalloc(&a0, nf * Ho * Ho * 192 * 2 + pad)); VG_CUDA(cudaMalloc(&s0, nf * Ho * Ho * 64 * 2 + pad));
  VG_CUDA(cudaMalloc(&x, mx * 2 + pad)); VG_CUDA(cudaMalloc(&y, mx * 2 + pad)); VG_CUDA(cudaMalloc(&idn, mx * 2 + pad));
  VG_CUDA(cudaMalloc(&o1, mo1 * 2 + pad)); VG_CUDA(cudaMalloc(&o2, mo2 * 2 + pad));
  VG_CUDA(cudaMalloc(&a2, std::max<size_t>(ma2, 1) * 2 + pad)); VG_CUDA(cudaMalloc(&xs, std::max<size_t>(mxs, 1) * 2 + pad));
  cap_frames = frames; cap_R = R;
}

void ResNet::release() {
  for (void* q : {(void*)a0, (void*)s0, (void*)x, (void*)y, (void*)idn, (void*)o1, (void*)o2, (void*)a2, (void*)xs})
    if (q) cudaFree(q);
  a0 = s0 = x = y = idn = o1 = o2 = a2 = xs = nullptr;
}

// ------------------------------------------------------------------------------------------------ forward
int ResNet::forward(const float* frames, int n, int R, bf16* out_bf16, float* out_f32, float* const* layer_out, cudaStream_t st) {
  VG_CHECK(loaded, "vgqa_resnet_backbone needs the 'vis_encoder.0.body.*' weights (conv1, bn1, layer1-4)");
  VG_CHECK(n >= 1 && R >= 32 && R % 32 == 0 && R <= 512, "vgqa_resnet_backbone: the frame side must be a multiple of 32 (at most 512)");
  const int launches0 = launches;
  // frames per pass (bounds the workspace: 2.7 GB at 224 px).  One frame is two 128-row tiles of the 14x14 (+ border) maps of
  // layer3, where most of the time goes: a pass of as many frames as there are SMs fills whole waves of its 256-column GEMMs
  int sms = 148;
  { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int chunk = std::min(n, std::max(sms, 32));
  ensure_workspace(chunk, R);
  const int Ho = R / 2;
  for (int f0 = 0; f0 < n; f0 += chunk) {
    const int nf = std::min(chunk, n - f0);
    {  // stem: conv 7x7 / 2 + FrozenBN + ReLU (GEMM over the staged operand), then MaxPool 3x3 / 2 into the padded grid of layer1
      stem_im2col_kernel<<<nf * Ho, 256, 21 * (R + 6) * 4, st>>>(frames + (size_t)f0 * 3 * R * R, a0, R);
      conv_gemm(a0, stem, 1, nf * Ho * Ho, 0, 0, nullptr, true, s0, st);
      const int Hp = Ho / 2 + 2;
      const long long total = (long long)nf * Hp * Hp * 8;
      maxpool_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const uint4*>(s0), reinterpret_cast<uint4*>(x), Ho, 8, total);
      launches += 3;
    }
    int H = R / 4, cin = 64;
    for (int l = 0; l < kLayers; ++l) {
      const int width = 64 << l;
      for (size_t b = 0; b < layer[l].size(); ++b) {
        const Block& k = layer[l][b];
        const int Hp_in = H + 2, rows_in = nf * Hp_in * Hp_in;
        const bf16* identity = x;
        conv_gemm(x, k.c1, 1, rows_in, Hp_in, Hp_in, nullptr, true, o1, st);   // 1x1 + ReLU at the input resolution
        ++launches;
        if (k.stride == 2) {
          const int Hq = H / 2, Hp = Hq + 2, rows = nf * Hp * Hp;
          long long total = (long long)rows * 9 * (width / 8);
          im2col3x3_s2_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const uint4*>(o1), reinterpret_cast<uint4*>(a2), H, width / 8, total);
          conv_gemm(a2, k.c2, 1, rows, Hp, Hp, nullptr, true, o2, st);           // 3x3 / 2 as a GEMM over the explicit operand
          total = (long long)rows * (cin / 8);
          subsample_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(xs), H, cin / 8, total);
          conv_gemm(xs, k.down, 1, rows, Hp, Hp, nullptr, false, idn, st);       // downsample: 1x1 / 2 + FrozenBN
          identity = idn;
          launches += 4;
          H = Hq;
        } else {
          conv_gemm(o1, k.c2, 9, rows_in, Hp_in, Hp_in, nullptr, true, o2, st);  // 3x3: nine shifted row blocks of the same operand
          ++launches;
          if (k.has_down) {
            conv_gemm(x, k.down, 1, rows_in, Hp_in, Hp_in, nullptr, false, idn, st);
            identity = idn;
            ++launches;
          }
        }
        const int Hp = H + 2, rows = nf * Hp * Hp;
        conv_gemm(o2, k.c3, 1, rows, Hp, Hp, identity, true, y, st);             // 1x1 + FrozenBN + identity → ReLU
        ++launches;
        std::swap(x, y);
        cin = 4 * width;
      }
      if (layer_out != nullptr && layer_out[l] != nullptr) {
        const long long total = (long long)nf * H * H * (cin / 8);
        unpad_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const uint4*>(x), nullptr,
                                                      reinterpret_cast<float4*>(layer_out[l] + (size_t)f0 * H * H * cin), H, cin / 8, total);
        ++launches;
      }
    }
    const long long total = (long long)nf * H * H * (2048 / 8);
    unpad_kernel<<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const uint4*>(x),
                                                  out_bf16 ? reinterpret_cast<uint4*>(out_bf16 + (size_t)f0 * H * H * 2048) : nullptr,
                                                  out_f32 ? reinterpret_cast<float4*>(out_f32 + (size_t)f0 * H * H * 2048) : nullptr, H, 256, total);
    ++launches;
    VG_CUDA(cudaGetLastError());
  }
  return launches - launches0;
}

}  // namespace vg
