// Fused post-norm FFN block of the cross-modal encoder (modal_encoder.py:175-177):
//
//   y = LayerNorm( x32 + W2 · relu(W1 · x + b1) + b2 )          x: [M,256] bf16, W1: [F,256], W2: [256,F], F % 128 == 0
//   C = bf16(y)    C32 = y (fp32 residual stream)    C2 = bf16(y + pos[row % period])   (next layer's Q/K operand)
//
// The two-kernel path (gemm_ws FFN1 + gemm_ln FFN2) writes and re-reads the [M,F] hidden activation: 8 KB of HBM traffic
// per row and layer, 42 % of the encoder's bytes.  Here the hidden activation never leaves the SM pair:
//
//   * one CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns a 256-row tile; each CTA holds its 128 rows of x (64 KB,
//     TMA, resident for the whole tile) and HALF of every weight block (the pair shares B operands through the
//     2-SM MMA, which halves the L2→SM weight stream: 2 MB of weights per 256 rows instead of per 128);
//   * the hidden dimension is walked in chunks of 128:   Hacc[c&1] (TMEM, 128 cols) = x · W1[c]^T      (16 UMMA 256x128x16)
//                                                         Hs[c&1]  (smem, bf16)      = relu(Hacc + b1)  (epilogue warps)
//                                                         Out      (TMEM, 256 cols) += Hs · W2[:,c]^T  ( 8 UMMA 256x256x16)
//     with the first GEMM running two chunks ahead of the second so the tensor pipe never waits for the ReLU pass;
//   * after the last chunk the epilogue warps pull Out into registers (releasing it at once for the next tile), add the
//     fp32 residual and b2, compute the LayerNorm moments with one exchange between the warps sharing a row, and write
//     the three outputs with coalesced 16-byte stores through a small swizzled slab.
//
// TMEM: Out 256 + Hacc 2 x 128 = 512 columns.  Shared memory per CTA: x 64 KB + Hs 2 x 32 KB + weight ring 4 x 16 KB +
// slabs 32 KB.  HBM traffic per row: 512 B (x) + 1 KB (residual) + 512 B + 1 KB + 512 B (outputs) = 3.5 KB.
#include "common.h"
#include "ptx.cuh"

namespace vg {

static constexpr int kFfnStages = 4;
static constexpr int kFfnStageBytes = 16384;
static constexpr int kFfnXBytes = 4 * 16384;
static constexpr int kFfnHBytes = 2 * 32768;
static constexpr int kFfnSlabBytes = 32768;
static constexpr int kFfnMiscBytes = 256 + 1024;  // barriers + TMEM slot, b1 chunk double buffer
static_assert(kFfnXBytes + kFfnHBytes + kFfnStages * kFfnStageBytes + kFfnSlabBytes + kFfnMiscBytes <= 232448, "smem");
static constexpr int kFfnSmem = kFfnXBytes + kFfnHBytes + kFfnStages * kFfnStageBytes + kFfnSlabBytes + kFfnMiscBytes;

struct FfnParams {
  const float* b1; const float* b2; const float* ln_w; const float* ln_b;
  const float* res32;   // [M,256] fp32 or nullptr
  float* C32;           // [M,256] fp32 or nullptr
  bf16* C;              // [M,256]
  bf16* C2;             // [M,256] or nullptr
  const bf16* add2;     // [period,256]
  int M, F, add2_period;
  float eps;
};

#ifdef VGQA_FFN_PROFILE
__device__ long long g_ffn_prof[2 * 16384];
#define FFN_PROF(slot, k) do { if (blockIdx.x < 2 && (slot) >= 0 && (slot) < 1024 && elect_one()) g_ffn_prof[blockIdx.x * 16384 + (slot) * 16 + (k)] = clock64(); } while (0)
#else
#define FFN_PROF(slot, k) do { } while (0)
#endif

// Warp roles (16 warps, registers re-balanced with setmaxnreg: 128 at launch → 56 / 104 / 176):
//   warp 0      TMA producer (weights; both CTAs)          warp 1      tcgen05.mma issuer (leader CTA only)
//   warps 4-7   hidden-activation pass (TMEM → ReLU → bf16 smem), one warp per TMEM lane quadrant
//   warps 8-15  LayerNorm pass, two warps per quadrant (128 output columns each, row values held in registers)
static constexpr int kFfnThreads = 512;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFfnThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w1,
                 const __grid_constant__ CUtensorMap tma_w2, const FfnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sx = smem;
  uint8_t* sh = sx + kFfnXBytes;
  uint8_t* sw = sh + kFfnHBytes;
  uint8_t* sslab = sw + kFfnStages * kFfnStageBytes;     // 8 x 4 KB (LayerNorm warps)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sslab + kFfnSlabBytes);
  uint64_t* full_bar = bars;                       // [stages]  leader: 1 arrival + tx bytes from both CTAs
  uint64_t* empty_bar = full_bar + kFfnStages;     // [stages]  both CTAs (multicast commit)
  uint64_t* x_full = empty_bar + kFfnStages;       // leader
  uint64_t* hacc_full = x_full + 1;                // [2] both (multicast commit)
  uint64_t* h_ready = hacc_full + 2;               // [2 buffers][2 k-blocks] leader, 8 arrivals each (4 warps x 2 CTAs)
  uint64_t* out_full = h_ready + 4;                // both (multicast commit)
  uint64_t* out_empty = out_full + 1;              // leader, 16 arrivals (8 warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_empty + 1);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2][128] b1 of the current / next chunk

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_tiles = (p.M + 255) / 256;
  const int n_my = (num_tiles - pair + npairs - 1) / npairs;
  const int nchunk = p.F / 128;

  if (warp == 0 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tma_x); tma_prefetch_desc(&tma_w1); tma_prefetch_desc(&tma_w2);
    for (int s = 0; s < kFfnStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(x_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&hacc_full[b], 1); mbar_init(&h_ready[2 * b], 8); mbar_init(&h_ready[2 * b + 1], 8); }
    mbar_init(out_full, 1);
    mbar_init(out_empty, 16);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_out = tmem_base, tm_h = tmem_base + 256;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ===================== TMA producer (both CTAs; each loads its half of every weight block) =====================
      if (lane == 0) {
        const uint32_t full_remote0 = mapa_u32(smem_u32(&full_bar[0]), 0);   // leader's full barriers (8 bytes apart)
        int stage = 0;
        uint32_t phase = 0;
        auto load_w1 = [&](int c) __attribute__((always_inline)) {   // W1 rows [c*128 + rank*64, +64): two stages of 2 x (64 x 64) boxes
          for (int s2 = 0; s2 < 2; ++s2) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * kFfnStageBytes);
            uint8_t* dst = sw + stage * kFfnStageBytes;
            tma_load_2d_pair(dst, &tma_w1, full_remote0 + 8u * stage, (2 * s2) * 64, c * 128 + (int)rank * 64);
            tma_load_2d_pair(dst + 8192, &tma_w1, full_remote0 + 8u * stage, (2 * s2 + 1) * 64, c * 128 + (int)rank * 64);
            if (++stage == kFfnStages) { stage = 0; phase ^= 1; }
          }
        };
        auto load_w2 = [&](int c) __attribute__((always_inline)) {   // W2 rows [rank*128, +128), k = hidden [c*128, +128): two 128 x 64 boxes
          for (int s2 = 0; s2 < 2; ++s2) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * kFfnStageBytes);
            tma_load_2d_pair(sw + stage * kFfnStageBytes, &tma_w2, full_remote0 + 8u * stage, c * 128 + s2 * 64, (int)rank * 128);
            if (++stage == kFfnStages) { stage = 0; phase ^= 1; }
          }
        };
        {  // x rows of the first tile (later tiles are fetched by the hidden-pass warp as soon as the buffer is free)
          const uint32_t xr = mapa_u32(smem_u32(x_full), 0);
          if (leader) mbar_expect_tx(x_full, 2 * kFfnXBytes);
          for (int kb = 0; kb < 4; ++kb) tma_load_2d_pair(sx + kb * 16384, &tma_x, xr, kb * 64, pair * 256 + (int)rank * 128);
        }
        load_w1(0);
        load_w1(1);
        for (int it = 0; it < n_my; ++it) {
          for (int c = 0; c < nchunk; ++c) {
            load_w2(c);
            if (c + 2 < nchunk) load_w1(c + 2);
            else if (c == nchunk - 1 && it + 1 < n_my) { load_w1(0); load_w1(1); }
          }
        }
      }
    } else if (warp == 1 && leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      constexpr uint32_t idesc1 = umma_idesc_bf16(256, 128);
      constexpr uint32_t idesc2 = umma_idesc_bf16(256, 256);
      int stage = 0;
      uint32_t phase = 0;
      auto gemm1 = [&](int g) __attribute__((always_inline)) {   // Hacc[g&1] = x · W1[chunk]^T
        const uint32_t d = tm_h + (g & 1) * 128;
        for (int s2 = 0; s2 < 2; ++s2) {
          mbar_wait(&full_bar[stage], phase);
          FFN_PROF(g - 2, 2 + s2);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kb2 = 0; kb2 < 2; ++kb2) {
              const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(sx + (2 * s2 + kb2) * 16384));
              const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(sw + stage * kFfnStageBytes + kb2 * 8192));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_pair(d, adesc + 2 * k, bdesc + 2 * k, idesc1, (s2 | kb2 | k) != 0 ? 1u : 0u);
            }
            umma_commit_pair_mc(&empty_bar[stage], 3);
            if (s2 == 1) umma_commit_pair_mc(&hacc_full[g & 1], 3);
          }
          __syncwarp();
          if (++stage == kFfnStages) { stage = 0; phase ^= 1; }
        }
      };
      auto gemm2 = [&](int g, int c, int it) __attribute__((always_inline)) {       // Out += Hs[g&1] · W2[:, chunk]^T
        if (c == 0) { mbar_wait(out_empty, (it & 1) ^ 1); tc_fence_after(); }
        for (int s2 = 0; s2 < 2; ++s2) {
          mbar_wait(&h_ready[(g & 1) * 2 + s2], (g >> 1) & 1);   // this 64-column half of the hidden chunk is in smem
          if (s2 == 0) FFN_PROF(g, 1);
          mbar_wait(&full_bar[stage], phase);
          FFN_PROF(g, 4 + s2);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(sh + (g & 1) * 32768 + s2 * 16384));
            const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(sw + stage * kFfnStageBytes));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_pair(tm_out, adesc + 2 * k, bdesc + 2 * k, idesc2, (c | s2 | k) != 0 ? 1u : 0u);
            umma_commit_pair_mc(&empty_bar[stage], 3);
            if (s2 == 1 && c == nchunk - 1) umma_commit_pair_mc(out_full, 3);
          }
          __syncwarp();
          if (++stage == kFfnStages) { stage = 0; phase ^= 1; }
        }
      };
      mbar_wait(x_full, 0);
      tc_fence_after();
      gemm1(0);
      gemm1(1);
      for (int it = 0; it < n_my; ++it) {
        for (int c = 0; c < nchunk; ++c) {
          const int g = it * nchunk + c;
          FFN_PROF(g, 0);
          gemm2(g, c, it);
          if (c + 2 < nchunk) {
            gemm1(g + 2);
          } else if (c == nchunk - 1 && it + 1 < n_my) {
            mbar_wait(x_full, (it + 1) & 1);
            tc_fence_after();
            gemm1(g + 1);
            gemm1(g + 2);
          }
        }
      }
    }
  } else if (warp < 8) {
    // ===================== hidden pass: Hs[b] = bf16(relu(Hacc[b] + b1)), one warp per lane quadrant =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    const int quad = warp & 3;
    const int row_l = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t h_ready_remote0 = mapa_u32(smem_u32(&h_ready[0]), 0);
    const uint32_t x_full_remote = mapa_u32(smem_u32(x_full), 0);
    const int swz = lane & 7;
    // b1 of the current chunk sits in shared memory (read with broadcast 16-byte loads); the four warps stage the next
    // chunk's 128 values one chunk ahead (one coalesced load per lane), published by a 128-thread named barrier.
    sbias[quad * 32 + lane] = __ldg(p.b1 + quad * 32 + lane);
    for (int it = 0; it < n_my; ++it) {
      for (int c = 0; c < nchunk; ++c) {
        const int g = it * nchunk + c, b = g & 1;
        const float bnext = __ldg(p.b1 + (c + 1 < nchunk ? c + 1 : 0) * 128 + quad * 32 + lane);
        mbar_wait(&hacc_full[b], (g >> 1) & 1);   // first GEMM of chunk g done (and, in issue order, the second GEMM of g-2:
        if (quad == 0) FFN_PROF(g, 6);            //  Hs[b] is free again)
        tc_fence_after();
        if (c == nchunk - 1 && it + 1 < n_my && quad == 0 && lane == 0) {
          // every first-GEMM of this tile has completed → the x buffer is free: fetch the next tile's rows
          const int tile_n = pair + (it + 1) * npairs;
          if (leader) mbar_expect_tx(x_full, 2 * kFfnXBytes);
          for (int kb = 0; kb < 4; ++kb) tma_load_2d_pair(sx + kb * 16384, &tma_x, x_full_remote, kb * 64, tile_n * 256 + (int)rank * 128);
        }
        named_bar_sync(5, 128);                    // chunk g-1 is finished by all four warps; sbias[g&1] is visible
        sbias[(b ^ 1) * 128 + quad * 32 + lane] = bnext;
        const float4* sb4 = reinterpret_cast<const float4*>(sbias + b * 128);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          uint32_t raw0[32], raw1[32];
          tmem_ld32(tm_h + lane_base + b * 128 + jj * 64, raw0);
          tmem_ld32(tm_h + lane_base + b * 128 + jj * 64 + 32, raw1);
          tmem_ld_wait();
          uint8_t* rowp = sh + b * 32768 + jj * 16384 + row_l * 128;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const uint32_t* raw = h2 ? raw1 : raw0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 ba = sb4[jj * 16 + h2 * 8 + 2 * u], bb = sb4[jj * 16 + h2 * 8 + 2 * u + 1];
              const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
              float a[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) a[i] = fmaxf(__uint_as_float(raw[8 * u + i]) + bv[i], 0.f);
              *reinterpret_cast<uint4*>(rowp + (((h2 * 4 + u) ^ swz) << 4)) =
                  make_uint4(pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]), pack_bf16(a[4], a[5]), pack_bf16(a[6], a[7]));
            }
          }
          if (jj == 1) tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(h_ready_remote0 + 8u * (b * 2 + jj));
          if (quad == 0) FFN_PROF(g, 7 + jj);
        }
      }
    }
  } else {
    // ===================== LayerNorm pass: warps 8..15, two per lane quadrant (128 output columns each) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    constexpr int CW = 128;
    const int ew = warp - 8;
    const int quad = warp & 3;
    const int part = ew >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t out_empty_remote = mapa_u32(smem_u32(out_empty), 0);
    uint8_t* slab = sslab + ew * 4096;
    uint8_t* slab_partner = sslab + (ew ^ 4) * 4096;
    const int swz = lane & 7;
    const int rl = lane >> 3, ul = lane & 7;   // read-back: row within a group of 4, 16-byte unit within the 128-byte slab row
    for (int it = 0; it < n_my; ++it) {
      const int tile = pair + it * npairs;
      const int row0 = tile * 256 + (int)rank * 128 + quad * 32;   // first row of this warp
      const int row = row0 + lane;
      const bool valid = row < p.M;
      float v[CW];
      // ---- drain: Out → registers, released at once so the next tile's second GEMM can start
      mbar_wait(out_full, it & 1);
      if (ew == 0) FFN_PROF(it * nchunk, 12);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < CW / 32; ++j) {
        uint32_t(&dst)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[j * 32]);
        tmem_ld32(tm_out + lane_base + part * CW + j * 32, dst);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(out_empty_remote);
      if (ew == 0) FFN_PROF(it * nchunk, 13);
      {  // ---- v += residual + b2.  The residual rows come in coalesced 64-byte row segments (cp.async) through the two
         //      halves of the slab.
        const bool has_res = p.res32 != nullptr;
        const int r4l = lane >> 2, u4 = lane & 3;
        auto issue = [&](int h) __attribute__((always_inline)) {   // h = 0..7: 16 fp32 columns [h*16, +16) of this warp's 128
          if (has_res) {
#pragma unroll
            for (int itr = 0; itr < 4; ++itr) {
              const int r = itr * 8 + r4l;
              const int u = (h & 1) * 4 + u4;
              if (row0 + r < p.M)
                cp_async16(slab + r * 128 + ((u ^ (r & 7)) << 4),
                           p.res32 + (size_t)(row0 + r) * 256 + part * CW + (h >> 1) * 32 + u * 4);
            }
          }
          cp_async_commit();
        };
        issue(0);
        issue(1);
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          if (h < 7) cp_async_wait<1>(); else cp_async_wait<0>();
          __syncwarp();
          const float4* b4 = reinterpret_cast<const float4*>(p.b2 + part * CW + h * 16);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 b = __ldg(b4 + u);
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_res && valid) r = *reinterpret_cast<const float4*>(slab + lane * 128 + ((((h & 1) * 4 + u) ^ swz) << 4));
            v[h * 16 + 4 * u] += b.x + r.x; v[h * 16 + 4 * u + 1] += b.y + r.y; v[h * 16 + 4 * u + 2] += b.z + r.z; v[h * 16 + 4 * u + 3] += b.w + r.w;
          }
          __syncwarp();
          if (h + 2 < 8) issue(h + 2);
        }
      }
      if (ew == 0) FFN_PROF(it * nchunk, 9);
      // ---- LayerNorm moments: the two warps sharing a row exchange (mean, M2) of their halves (parallel-variance combine)
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < CW; ++i) s += v[i];
      const float mean_p = s * (1.0f / CW);
      float m2_p = 0.f;
#pragma unroll
      for (int i = 0; i < CW; ++i) { const float d = v[i] - mean_p; m2_p = fmaf(d, d, m2_p); }
      reinterpret_cast<float2*>(slab)[lane] = make_float2(mean_p, m2_p);
      named_bar_sync(1 + quad, 64);
      const float2 o = reinterpret_cast<const float2*>(slab_partner)[lane];
      const float mean = 0.5f * (mean_p + o.x);
      const float dm = mean_p - o.x;
      const float rstd = rsqrtf((m2_p + o.y + 0.5f * (float)CW * dm * dm) * (1.0f / 256.0f) + p.eps);
      named_bar_sync(1 + quad, 64);   // the partner has read the exchange area before the slab is reused
      if (ew == 0) FFN_PROF(it * nchunk, 10);
      // ---- normalise → slab (32 rows x 32 fp32) → coalesced 16-byte stores of the three outputs
      int prow_base = 0;
      if (p.C2) prow_base = row0 % p.add2_period;
#pragma unroll
      for (int j = 0; j < CW / 32; ++j) {
        {
          uint8_t* rowp = slab + lane * 128;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i0 = j * 32 + u * 4;
            float4 z;
            z.x = (v[i0] - mean) * rstd; z.y = (v[i0 + 1] - mean) * rstd; z.z = (v[i0 + 2] - mean) * rstd; z.w = (v[i0 + 3] - mean) * rstd;
            *reinterpret_cast<float4*>(rowp + ((u ^ swz) << 4)) = z;
          }
        }
        __syncwarp();
        const int col = part * CW + j * 32 + ul * 4;
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.ln_w + col));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col));
        uint2 posv[8];
        if (p.C2) {
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            int pr = prow_base + itr * 4 + rl;
            while (pr >= p.add2_period) pr -= p.add2_period;
            posv[itr] = __ldg(reinterpret_cast<const uint2*>(p.add2 + (size_t)pr * 256 + col));
          }
        }
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {
          const int r = itr * 4 + rl;
          const float4 z = *reinterpret_cast<const float4*>(slab + r * 128 + ((ul ^ (r & 7)) << 4));
          const int grow = row0 + r;
          if (grow < p.M) {
            const float y0 = fmaf(z.x, w4.x, b4.x), y1 = fmaf(z.y, w4.y, b4.y), y2 = fmaf(z.z, w4.z, b4.z), y3 = fmaf(z.w, w4.w, b4.w);
            const size_t o2 = (size_t)grow * 256 + col;
            if (p.C32) __stcs(reinterpret_cast<float4*>(p.C32 + o2), make_float4(y0, y1, y2, y3));
            __stcs(reinterpret_cast<uint2*>(p.C + o2), make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3)));
            if (p.C2) {
              const float2 a = unpack_bf16(posv[itr].x), b = unpack_bf16(posv[itr].y);
              __stcs(reinterpret_cast<uint2*>(p.C2 + o2), make_uint2(pack_bf16(y0 + a.x, y1 + a.y), pack_bf16(y2 + b.x, y3 + b.y)));
            }
          }
        }
        __syncwarp();
        if (j == 0 && ew == 0) FFN_PROF(it * nchunk, 11);
      }
      if (ew == 0) FFN_PROF(it * nchunk, 14);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------ host
CUtensorMap make_tmap_2d(const void* ptr, int rows, int cols, int ld, int box_rows, bool f32);  // gemm_tc.cu
int device_sm_count();
void count_gemm_launch();

#ifdef VGQA_FFN_PROFILE
void ffn_prof_read(long long* dst) { cudaMemcpyFromSymbol(dst, g_ffn_prof, sizeof(long long) * 32768); }
#else
void ffn_prof_read(long long* dst) { for (int i = 0; i < 32768; ++i) dst[i] = 0; }
#endif

bool ffn_fused_supported(int F) { return F % 128 == 0 && F >= 256; }

static void launch_ffn(const CUtensorMap& tx, const CUtensorMap& tw1, const CUtensorMap& tw2, const FfnParams& p,
                       cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VG_CUDA(cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnSmem));
    attr_set = true;
  }
  // one CTA pair per TPC (pairs beyond what is co-resident simply start later: the tile loop is a static stride);
  // device_sm_count() honours the model's SM budget
  const int tiles = (p.M + 255) / 256;
  int pairs = device_sm_count() / 2;
  if (pairs > tiles) pairs = tiles;
  ffn_fused_kernel<<<2 * pairs, kFfnThreads, kFfnSmem, stream>>>(tx, tw1, tw2, p);
  VG_CUDA(cudaGetLastError());
  count_gemm_launch();
}

// y = LN(res32 + W2 relu(W1 x + b1) + b2); see the header comment.  x: [M,256] bf16 (ld 256); all outputs ld 256.
void ffn_fused(const bf16* X, const bf16* W1, const float* b1, const bf16* W2, const float* b2, int M, int F,
               const float* res32, const float* ln_w, const float* ln_b, float eps, bf16* C, float* C32, bf16* C2,
               const bf16* add2, int add2_period, int epi_parts, cudaStream_t stream) {
  VG_CHECK(ffn_fused_supported(F), "ffn_fused: hidden size must be a multiple of 128");
  VG_CHECK(M > 0 && C != nullptr && b1 && b2 && ln_w && ln_b, "ffn_fused: bad arguments");
  VG_CHECK(C2 == nullptr || add2 != nullptr, "ffn_fused: C2 needs add2");
  FfnParams p;
  p.b1 = b1; p.b2 = b2; p.ln_w = ln_w; p.ln_b = ln_b; p.res32 = res32; p.C32 = C32; p.C = C; p.C2 = C2; p.add2 = add2;
  p.M = M; p.F = F; p.add2_period = add2_period > 0 ? add2_period : 1; p.eps = eps;
  CUtensorMap tx = make_tmap_2d(X, M, 256, 256, 128, false);
  CUtensorMap tw1 = make_tmap_2d(W1, F, 256, 256, 64, false);
  CUtensorMap tw2 = make_tmap_2d(W2, 256, F, F, 128, false);
  (void)epi_parts;
  launch_ffn(tx, tw1, tw2, p, stream);
}

}  // namespace vg
