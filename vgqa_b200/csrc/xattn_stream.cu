// xattn1, streaming form: the absorbed one-query cross-attention of the decoders / SpatialActivation
//
//   s[h, m]  = q~[f, h, :] · mem[f, m, :]  (+ sbias[f, h, m])          h < 8 heads, m < Mk memory tokens of frame f
//   p[h, :]  = softmax_m(scale · s[h, :])                               (optional key padding mask)
//   ctx[f, h, :] = sum_m p[h, m] · mem[f, m, :]                         [F, 8*256] bf16
//   att[f, m] = minmax_m(sigmoid(sum_h p[h, m]))                        (optional; classifier.py:72-78)
//
// The op is a GEMV-like 16 FLOP/B stream over the memory tokens: HBM-bound.  The CTA-per-frame kernel in attention.cu
// is a load → sync → compute → sync → store chain whose throughput comes only from 4 resident CTAs per SM (38–52 us for
// 135–177 MB).  Here a PAIR of warps is an independent worker with a private TMA ring (memory tokens of a frame as four
// 128B-swizzled column blocks + the 8 absorbed query rows), so the next frame's 40 KB are in flight while the current frame
// is processed; the two warps split the keys (scores) and the channels (context) and meet at two named barriers per frame:
//   phase 1   S[key, head] = mem · q~^T with mma.sync m16n8k16: keys are the M index (16 per tile, no padding), the 8 heads
//             the N index; A = mem rows via ldmatrix, B = q~ rows via ldmatrix
//   softmax   in the accumulator registers (a head's keys are spread over the 8 row groups → three shuffles per
//             reduction); the normalised probabilities go through a 1.4 KB transposed tile Ps[head][key]
//   phase 2   ctx^T[channel, head] = mem^T · P: channels are the M index (A = mem via ldmatrix.trans), B = Ps via ldmatrix;
//             results staged over the consumed query tile → one TMA store.
// 16 MT + 16 MT mma per frame (MT = 16-key tiles), none of them padded.
#include <stdlib.h>
#include <string>

#include "common.h"
#include "ptx.cuh"

namespace vg {

struct XsParams {
  const float* sbias;    // optional [F, 8, ldsb] additive score term (unscaled)
  const uint8_t* kmask;  // optional [F, ldmask]
  float* att;            // optional [F, Mk]
  int F, Mk, ldsb, ldmask;
  float scale_log2e;
};

__device__ __forceinline__ void tma_load_3d_xs(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// MT = number of 16-key blocks held per frame (TMA box rows = 16*MT >= Mk); NP = worker PAIRS per CTA; NS = ring stages.
// A worker is a pair of warps sharing one frame: warp h takes the key tiles jb = h, h+2, … in phase 1 and the channel tiles
// [8h, 8h+8) in phase 2; the two meet at two 64-thread named barriers per frame (softmax merge, probabilities complete).
template <int MT, int NP, int NS>
__global__ void __launch_bounds__(NP * 64, 1)
xattn_stream_kernel(const __grid_constant__ CUtensorMap tm_mem, const __grid_constant__ CUtensorMap tm_q,
                    const __grid_constant__ CUtensorMap tm_ctx, const XsParams p) {
  constexpr int MP = MT * 16;
  constexpr int kMemBytes = 4 * MP * 128;          // four column blocks of MP rows x 128 B
  constexpr int kStageBytes = kMemBytes + 4096;    // + absorbed queries: four blocks of 8 rows x 128 B
  constexpr int kPsRow = (MP + 8) * 2;             // bytes per head row of the transposed probabilities (padded: conflict-free ldmatrix)
  constexpr int NCS = MT >= 13 ? 1 : 2;            // ctx staging tiles (one when shared memory is tight)
  constexpr int kXchBytes = 2 * 8 * 16;            // per warp of the pair: [8 heads] x (max, sum, att min, att max)
  constexpr int kPairBytes = NS * kStageBytes + NCS * 4096 + 8 * kPsRow + kXchBytes;
  constexpr int MTH = (MT + 1) / 2;                // key tiles per warp
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, h = warp & 1;
  uint8_t* wbase = smem + (size_t)pair * NS * kStageBytes;
  uint8_t* cstage = smem + (size_t)NP * NS * kStageBytes + pair * (NCS * 4096);
  uint8_t* ps = smem + (size_t)NP * (NS * kStageBytes + NCS * 4096) + pair * (8 * kPsRow);
  float4* xch = reinterpret_cast<float4*>(smem + (size_t)NP * (NS * kStageBytes + NCS * 4096 + 8 * kPsRow) + pair * kXchBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NP * kPairBytes) + pair * 2 * NS;
  uint64_t* sfree = full + NS;   // both warps of the pair are done reading a stage
  const int gid = lane >> 2, tq = lane & 3;
  const int worker = blockIdx.x * NP + pair, nworkers = gridDim.x * NP;
  const int Mk = p.Mk;
  const bool issuer = h == 0 && lane == 0;

  if (issuer) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&sfree[i], 2); }
    fence_mbar_init();
    tma_prefetch_desc(&tm_mem); tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_ctx);
  }
  named_bar_sync(1 + pair, 64);

  auto issue = [&](int f, int s) {   // one thread of the pair: fetch frame f into stage s
    uint8_t* st = wbase + s * kStageBytes;
    mbar_expect_tx(&full[s], kStageBytes);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      tma_load_3d_xs(st + b * (MP * 128), &tm_mem, &full[s], b * 64, 0, f);
      tma_load_2d(st + kMemBytes + b * 1024, &tm_q, &full[s], b * 64, f * 8);
    }
  };

  int it = 0;
  if (issuer)
    for (int k = 0; k < NS - 1; ++k)
      if (worker + k * nworkers < p.F) issue(worker + k * nworkers, k);
  for (int f = worker; f < p.F; f += nworkers, ++it) {
    const int s = it % NS;
    if (issuer && f + (NS - 1) * nworkers < p.F) {
      // keep NS-1 frames in flight: refill the stage of the previous frame as soon as both warps have left it
      if (it > 0) mbar_wait(&sfree[(it - 1) % NS], ((it - 1) / NS) & 1);
      issue(f + (NS - 1) * nworkers, (it + NS - 1) % NS);
    }
    // Score accumulators sc[i] = S[key 16 jb + gid (+8)][head 2 tq (+1)] for this warp's tiles jb = h + 2 i, started from the
    // additive score term of this frame (fetched before the wait so the loads overlap the TMA transfer).
    float sc[MTH][4];
#pragma unroll
    for (int i = 0; i < MTH; ++i) sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
    if (p.sbias != nullptr) {
      const float* g = p.sbias + ((size_t)f * 8 + tq * 2) * p.ldsb;
#pragma unroll
      for (int i = 0; i < MTH; ++i) {
        const int k0 = (h + 2 * i) * 16 + gid, k1 = k0 + 8;
        if (k0 < Mk) { sc[i][0] = __ldg(g + k0); sc[i][1] = __ldg(g + p.ldsb + k0); }
        if (k1 < Mk) { sc[i][2] = __ldg(g + k1); sc[i][3] = __ldg(g + p.ldsb + k1); }
      }
    }
    mbar_wait(&full[s], (it / NS) & 1);
    const uint32_t mbase = smem_u32(wbase + s * kStageBytes);
    const uint32_t qbase = mbase + kMemBytes;

    // ---------------- phase 1: S[key][head] = mem · q~^T  (M = 16 keys per tile, N = 8 heads, K = 256 channels)
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      uint32_t bq[2];
      {  // B = q~^T: heads as the n index, k-columns [16 ks, +16) as two 8x8 matrices
        const int r = lane & 7, unit = ks * 2 + ((lane >> 3) & 1);
        ldmatrix_x2(bq, qbase + (unit >> 3) * 1024 + r * 128 + (((unit & 7) ^ r) << 4));
      }
#pragma unroll
      for (int i = 0; i < MTH; ++i) {   // A: (keys 0-7, k lo), (keys 8-15, k lo), (keys 0-7, k hi), (keys 8-15, k hi)
        const int jb = h + 2 * i;
        if (jb < MT) {
          const int m = lane >> 3;
          const int row = jb * 16 + ((m & 1) << 3) + (lane & 7);
          const int unit = ks * 2 + (m >> 1);
          uint32_t a[4];
          ldmatrix_x4(a, mbase + (unit >> 3) * (MP * 128) + row * 128 + (((unit & 7) ^ (row & 7)) << 4));
          mma_16816(sc[i], a, bq);
        }
      }
    }
    // ---------------- softmax over the keys (rows), two heads per thread; the pair merges its two key sets flash-style
    const uint8_t* km = p.kmask ? p.kmask + (size_t)f * p.ldmask : nullptr;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < MTH; ++i) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int key = (h + 2 * i) * 16 + gid + hf * 8;
        const bool dead = key >= Mk || (km != nullptr && km[key] != 0);
        const float v0 = dead ? -INFINITY : sc[i][2 * hf] * p.scale_log2e;
        const float v1 = dead ? -INFINITY : sc[i][2 * hf + 1] * p.scale_log2e;
        sc[i][2 * hf] = v0; sc[i][2 * hf + 1] = v1;
        mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1);
      }
    }
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
    }
    if (mx0 == -INFINITY) mx0 = 0.f;
    if (mx1 == -INFINITY) mx1 = 0.f;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int i = 0; i < MTH; ++i) {
      sc[i][0] = exp2f(sc[i][0] - mx0); sc[i][2] = exp2f(sc[i][2] - mx0);
      sc[i][1] = exp2f(sc[i][1] - mx1); sc[i][3] = exp2f(sc[i][3] - mx1);
      sum0 += sc[i][0] + sc[i][2];
      sum1 += sc[i][1] + sc[i][3];
    }
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, o);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, o);
    }
    if (gid == 0) {   // lanes 0..3 publish (max, sum) of heads 2 tq, 2 tq + 1 over this warp's keys
      *reinterpret_cast<float2*>(&xch[h * 8 + tq * 2]) = make_float2(mx0, sum0);
      *reinterpret_cast<float2*>(&xch[h * 8 + tq * 2 + 1]) = make_float2(mx1, sum1);
    }
    named_bar_sync(1 + pair, 64);   // the partial softmax terms of both warps are visible
    float r0, r1;
    {
      const float4 o0 = xch[(h ^ 1) * 8 + tq * 2], o1 = xch[(h ^ 1) * 8 + tq * 2 + 1];
      const float m0 = fmaxf(mx0, o0.x), m1 = fmaxf(mx1, o1.x);
      const float t0 = sum0 * exp2f(mx0 - m0) + o0.y * exp2f(o0.x - m0);
      const float t1 = sum1 * exp2f(mx1 - m1) + o1.y * exp2f(o1.x - m1);
      r0 = exp2f(mx0 - m0) / t0;
      r1 = exp2f(mx1 - m1) / t1;
    }
    // normalised probabilities → transposed bf16 tile Ps[head][key] (the B operand of phase 2)
#pragma unroll
    for (int i = 0; i < MTH; ++i) {
      sc[i][0] *= r0; sc[i][2] *= r0; sc[i][1] *= r1; sc[i][3] *= r1;
      if (h + 2 * i < MT) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int key = (h + 2 * i) * 16 + gid + hf * 8;
          *reinterpret_cast<bf16*>(ps + (tq * 2) * kPsRow + key * 2) = __float2bfloat16(sc[i][2 * hf]);
          *reinterpret_cast<bf16*>(ps + (tq * 2 + 1) * kPsRow + key * 2) = __float2bfloat16(sc[i][2 * hf + 1]);
        }
      }
    }
    // ---------------- optional attention map: minmax(sigmoid(sum over heads)); the min/max spans both warps' keys
    float amin = INFINITY, amax = -INFINITY;
    if (p.att != nullptr) {
#pragma unroll
      for (int i = 0; i < MTH; ++i) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float a = sc[i][2 * hf] + sc[i][2 * hf + 1];
          a += __shfl_xor_sync(0xffffffffu, a, 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2);
          a = 1.f / (1.f + __expf(-a));
          sc[i][2 * hf] = a;
          if ((h + 2 * i) * 16 + gid + hf * 8 < Mk) { amin = fminf(amin, a); amax = fmaxf(amax, a); }
        }
      }
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        amin = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, o));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      }
      if (lane == 0) { xch[h * 8].z = amin; xch[h * 8].w = amax; }
    }
    named_bar_sync(1 + pair, 64);   // Ps (and the att extrema) of both warps are complete
    if (p.att != nullptr) {
      const float4 o = xch[(h ^ 1) * 8];
      amin = fminf(amin, o.z); amax = fmaxf(amax, o.w);
      const float ia = 1.f / (amax - amin + 1e-6f);
      if (tq == 0) {
#pragma unroll
        for (int i = 0; i < MTH; ++i) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int key = (h + 2 * i) * 16 + gid + hf * 8;
            if (key < Mk) p.att[(size_t)f * Mk + key] = (sc[i][2 * hf] - amin) * ia;
          }
        }
      }
    }
    // ---------------- phase 2: ctx^T[ch][head] = mem^T · P for this warp's 128 channels (M = 16 channels per tile, K = keys)
    uint32_t pb[MT][2];
#pragma unroll
    for (int kk = 0; kk < MT; ++kk)
      ldmatrix_x2(pb[kk], smem_u32(ps) + (lane & 7) * kPsRow + (kk * 16 + (((lane >> 3) & 1) << 3)) * 2);
    float d[8][4];
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) d[mt][0] = d[mt][1] = d[mt][2] = d[mt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < MT; ++kk) {
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) {   // A = mem^T via ldmatrix.trans: (keys lo, ch lo), (keys lo, ch hi), (keys hi, ch lo), (keys hi, ch hi)
        const int m = lane >> 3;
        const int row = kk * 16 + ((m >> 1) << 3) + (lane & 7);
        const int unit = (h * 8 + mt) * 2 + (m & 1);
        uint32_t a[4];
        ldmatrix_x4_trans(a, mbase + (unit >> 3) * (MP * 128) + row * 128 + (((unit & 7) ^ (row & 7)) << 4));
        mma_16816(d[mt], a, pb[kk]);
      }
    }
    uint8_t* qt = cstage + (NCS == 2 ? (it & 1) * 4096 : 0);
    if (lane == 0) tma_store_wait_read<NCS - 1>();   // this warp's earlier store out of this staging tile has drained it
    __syncwarp();
    {  // ctx[head 2 tq (+1)][ch 128 h + 16 mt + gid (+8)] → swizzled staging tile → TMA store of this warp's two column blocks
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int c = h * 128 + mt * 16 + gid + hf * 8;
          const int unit = (c >> 3) & 7;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int hd = tq * 2 + e;
            *reinterpret_cast<bf16*>(qt + (c >> 6) * 1024 + hd * 128 + ((unit ^ hd) << 4) + (c & 7) * 2) =
                __float2bfloat16(d[mt][2 * hf + e]);
          }
        }
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&sfree[s]);   // this warp no longer reads the frame's stage
      tma_store_2d(&tm_ctx, qt + (2 * h) * 1024, (2 * h) * 64, f * 8);
      tma_store_2d(&tm_ctx, qt + (2 * h + 1) * 1024, (2 * h + 1) * 64, f * 8);
      tma_store_commit();
    }
  }
  if (lane == 0) tma_store_wait_all<0>();
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled_xs)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
void* tensor_map_encode_fn();  // gemm_tc.cu
CUtensorMap make_tmap_2d(const void* ptr, int rows, int cols, int ld, int box_rows, bool f32);
int device_sm_count();

// [F frames][Mk tokens][256] bf16, token stride 512 B, frame stride `frame_stride_rows` rows; box = 64 cols x box_rows x 1.
static CUtensorMap make_tmap_mem(const bf16* ptr, int F, int Mk, long long frame_stride_rows, int box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[3] = {256, (cuuint64_t)Mk, (cuuint64_t)F};
  cuuint64_t gstr[2] = {512, (cuuint64_t)frame_stride_rows * 512};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<PFN_encodeTiled_xs>(tensor_map_encode_fn())(
      &m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (xattn memory) failed with CUresult " + std::to_string((int)r));
  return m;
}

template <int MT, int NP, int NS>
static void launch_xs(const CUtensorMap& tm, const CUtensorMap& tq, const CUtensorMap& tc, const XsParams& p, cudaStream_t st) {
  constexpr int smem = NP * (NS * (4 * MT * 16 * 128 + 4096) + (MT >= 13 ? 1 : 2) * 4096 + 8 * (MT * 16 + 8) * 2 + 256) + NP * NS * 16 + 1024;
  static_assert(smem <= 232448, "xattn_stream: shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    VG_CUDA(cudaFuncSetAttribute(xattn_stream_kernel<MT, NP, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  int grid = device_sm_count();
  const int need = (p.F + NP - 1) / NP;
  if (grid > need) grid = need;
  xattn_stream_kernel<MT, NP, NS><<<grid, NP * 64, smem, st>>>(tm, tq, tc, p);
  VG_CUDA(cudaGetLastError());
}

// frame-invariant positional terms only (they arrive as `sbias`); frames whose memory rows are 512-byte strided
bool xattn_stream_supported(int Mk, long long frame_stride_rows) {
  return Mk >= 1 && Mk <= 208 && frame_stride_rows >= Mk;
}

void xattn_stream(const bf16* qt, const bf16* mem, long long frame_stride_rows, int F, int Mk, const float* sbias, int ldsb,
                  const uint8_t* kmask, int ldmask, float scale, bf16* ctx, float* att, cudaStream_t stream) {
  VG_CHECK(F > 0 && xattn_stream_supported(Mk, frame_stride_rows), "xattn_stream: unsupported shape");
  const int MT = Mk <= 64 ? 4 : Mk <= 80 ? 5 : Mk <= 128 ? 8 : 13;
  XsParams p;
  p.sbias = sbias; p.kmask = kmask; p.att = att; p.F = F; p.Mk = Mk; p.ldsb = ldsb; p.ldmask = ldmask;
  p.scale_log2e = scale * 1.4426950408889634f;
  CUtensorMap tm = make_tmap_mem(mem, F, Mk, frame_stride_rows, MT * 16);
  CUtensorMap tq = make_tmap_2d(qt, F * 8, 256, 256, 8, false);
  CUtensorMap tc = make_tmap_2d(ctx, F * 8, 256, 256, 8, false);
  // Four single-stage worker pairs per SM (default) instead of two double-buffered ones: the kernel is bound by the per-frame
  // dependent chain, not by any unit (profiles/r01_xattn_stream_ncu.md), so frames in flight per SM are what counts:
  // 40.3 us vs 49.1 us at the decoder shape.  VGQA_XS_PAIRS=2 selects the double-buffered form (A/B runs).
  static const int pairs4 = [] { const char* e = getenv("VGQA_XS_PAIRS"); return e == nullptr || e[0] != '2' ? 1 : 0; }();
  if (pairs4 && MT == 4) { launch_xs<4, 4, 1>(tm, tq, tc, p, stream); return; }
  if (pairs4 && MT == 5) { launch_xs<5, 4, 1>(tm, tq, tc, p, stream); return; }
  if (pairs4 && MT == 8) { launch_xs<8, 2, 1>(tm, tq, tc, p, stream); return; }   // 65..128 memory tokens: two single-stage pairs
  switch (MT) {
    case 4: launch_xs<4, 2, 2>(tm, tq, tc, p, stream); break;
    case 5: launch_xs<5, 2, 2>(tm, tq, tc, p, stream); break;
    case 8: launch_xs<8, 1, 3>(tm, tq, tc, p, stream); break;
    default: launch_xs<13, 1, 2>(tm, tq, tc, p, stream); break;
  }
}

}  // namespace vg
