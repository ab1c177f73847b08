// input_proj / input_proj2 of VSTGNet (grounding_net.py:62,71,101,105): the 1x1 Conv2d that maps the extractor feature maps
// (ResNet101 layer4: 2048 channels, Video-Swin stage 3: 768 channels) to the 256-d tokens of the cross-modal encoder —
// SURVEY.md §8(f) rank 2, the step immediately before the hot path.
//
//   X[(f*S + tok0 + p), n] = sum_c W[n, c] * in[f, c, p] + b[n]        in: [F, C, P] fp32 (NCHW), W: [256, C] bf16
//
// writes the three tensors the first encoder layer reads (bf16 x, fp32 residual stream, bf16 x + pos) straight into their
// token-major rows, so the projected NCHW feature map never exists in HBM (the two-step path wrote 256·P fp32 per frame and
// read it back through nchw_to_tokens).
//
// The op reads 4·C bytes per token and does 512·C FLOP on it: 128 FLOP/B, below the B200 ridge (≈210 FLOP/B), i.e. it is bound
// by the HBM read of the fp32 features.  The features arrive channel-major ([c][p], p contiguous), the tensor core wants a
// K-major bf16 operand ([p][c]) — the transpose + conversion happens on the way into shared memory:
//
//   warp 0       TMA producer of the W k-blocks (256 x 64 bf16, 128B swizzle)
//   warp 2       feature producer.  kBulk (whole frames per tile, i.e. P <= 128, 16-byte aligned input): cp.async.bulk of each
//                frame's [64 channels x P tokens] fp32 chunk — ONE contiguous 256·P-byte range — into a 3-stage shared-memory
//                staging ring, so the features reach the SM through the TMA engine (no LSU wavefronts, no sector over-fetch
//                of the unaligned 49-float rows).  Measured: 218 us vs 232 us (768 channels), 449 us vs 444 us (2048 channels)
//                against the LDG form below — both forms sit at ≈4.6 TB/s (profiles/r01_input_proj_ncu.md).
//                Otherwise (frames sliced into 128-row tiles): an L2 prefetcher (cp.async.bulk.prefetch.L2, 8 k-blocks ahead)
//                and the converter warps read global memory with coalesced 4-byte loads.
//   warp 1       tcgen05.mma issuer (UMMA 128x256x16), two TMEM accumulators so the epilogue of tile i overlaps tile i+1
//   warps 4-7    epilogue: TMEM row per thread → swizzled 4 KB slab per warp → + bias → X32 / X / XP with every store
//                instruction covering four 128-byte (fp32) / 64-byte (bf16) row segments (row-per-thread stores cost 32 L1
//                wavefronts each and made the LSU data pipe the limiter: ncu 78 % → profiles/r01_input_proj_ncu.md)
//   warps 8-31   converters, three groups of 8 warps; group g owns the k-blocks with (k-block index % 3) == g (= staging stage g),
//                so three k-blocks are in flight per SM.  Thread = (token row, 32 of the 64 channels of the k-block): 32
//                conflict-free 4-byte reads (a warp reads 32 consecutive tokens of one channel) from the staging stage — or from
//                global memory in the fallback —, packed to bf16 and written as four 16-byte stores into the 128B-swizzled
//                K-major tile the UMMA descriptor expects.  Registers are re-balanced with setmaxnreg.
//
// A tile is 128 token rows: floor(128 / P) whole frames when a frame has P <= 128 tokens (7x7: two frames, 98 rows), or one
// 128-row slice of a frame otherwise (14x14: two slices).
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace vg {

static constexpr int kIpABytes = 128 * 64 * 2;
static constexpr int kIpBBytes = 256 * 64 * 2;
static constexpr int kIpStgBytes = 128 * 64 * 4;           // fp32 staging stage: up to 128 token rows x 64 channels
static constexpr int kIpSlabBytes = 4 * 4096;               // epilogue staging: 32 rows x 128 B per warp
static constexpr int kIpPrefetchDist = 8;                  // k-blocks the L2 prefetcher runs ahead of the MMA issuer
static constexpr int kIpGroups = 3;                        // converter groups (k-blocks in flight)
static constexpr int kIpThreads = 256 + kIpGroups * 256;  // 4 control warps + 4 epilogue warps + 8 warps per group

// kMode 0: NCHW fp32 input, converters read global memory (LDG form);  1: NCHW fp32 input through the bulk-copy staging ring;
// 2: NHWC bf16 input (channels-last, what a bf16 channels_last backbone emits): the map IS the K-major A operand → plain TMA tiles
enum { kIpLdg = 0, kIpBulk = 1, kIpNhwc = 2 };
template <int kMode>
struct IpCfg {
  static constexpr bool kBulk = kMode == kIpBulk;
  static constexpr int kWStages = kBulk ? 2 : 4;
  static constexpr int kAStages = kBulk ? 3 : 4;
  static constexpr int kStgStages = kBulk ? kIpGroups : 0;   // one staging stage per converter group
  static constexpr int kThreads = kMode == kIpNhwc ? 256 : kIpThreads;   // no converter warps when the input is already bf16 rows
  static constexpr int kSmem = kWStages * kIpBBytes + kAStages * kIpABytes + kStgStages * kIpStgBytes + kIpSlabBytes +
                               1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(kSmem <= 232448, "input_proj shared memory");
};

struct IpParams {
  const float* in;     // [F, C, P]
  const float* bias;   // [256]
  const bf16* pos;     // token-major positional rows [pos_frames * S, 256] (XP = x + pos) or nullptr when XP is nullptr
  bf16* X;             // [F*S, 256]
  float* X32;          // [F*S, 256] or nullptr
  bf16* XP;            // [F*S, 256] or nullptr
  int F, C, P, S, tok0;
  int pos_per_frame;   // 1: pos row = output row; 0: pos row = tok0 + p (one table shared by every frame)
  int fpt;             // frames per tile (P <= 128), else 0
  int tpf;             // 128-row slices per frame (P > 128); 0 in the NHWC form: a tile is 128 consecutive rows of [F*P, C]
  int num_tiles;
  int l2_prefetch;     // LDG form: `in` is 16-byte aligned → bulk L2 prefetch of the chunks ahead
};

// tile row r → (frame, token); false for the padding rows of a tile
__device__ __forceinline__ bool ip_row(const IpParams& p, int tile, int r, int& frame, int& tok) {
  if (p.fpt > 0) {
    const int fl = r / p.P;
    frame = tile * p.fpt + fl;
    tok = r - fl * p.P;
    return fl < p.fpt && frame < p.F;
  }
  if (p.tpf == 0) {   // flat rows (NHWC input)
    const int m = tile * 128 + r;
    frame = m / p.P;
    tok = m - frame * p.P;
    return frame < p.F;
  }
  frame = tile / p.tpf;
  tok = (tile - frame * p.tpf) * 128 + r;
  return tok < p.P;
}

__device__ __forceinline__ void bulk_prefetch_l2(const void* gptr, uint32_t bytes) {   // 16-byte aligned address and size
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
// contiguous global → shared copy through the TMA engine; completion bytes are posted to `bar` (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gptr, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gptr), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int kMode>
__global__ void __launch_bounds__(kIpThreads, 1)
input_proj_kernel(const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_a, const IpParams p) {
  using Cfg = IpCfg<kMode>;
  constexpr bool kBulk = kMode == kIpBulk;
  constexpr int kWS = Cfg::kWStages, kAS = Cfg::kAStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + kAS * kIpABytes;
  uint8_t* smem_stg = smem_b + kWS * kIpBBytes;
  uint8_t* smem_slab = smem_stg + Cfg::kStgStages * kIpStgBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_slab + kIpSlabBytes);
  uint64_t* w_full = bars;              // [4] TMA bytes of the W k-block
  uint64_t* w_empty = bars + 4;         // [4] tcgen05.commit
  uint64_t* a_full = bars + 8;          // [4] 8 arrivals (the converter warps of one group)
  uint64_t* a_empty = bars + 12;        // [4] tcgen05.commit
  uint64_t* stg_full = bars + 16;       // [3] bulk-copy bytes of the fp32 chunk(s)
  uint64_t* stg_empty = bars + 19;      // [3] 8 arrivals (the group has read the stage)
  uint64_t* tfull_bar = bars + 22;      // [2]
  uint64_t* tempty_bar = bars + 24;     // [2] 4 arrivals (epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  volatile int* consumed = reinterpret_cast<volatile int*>(tmem_slot + 1);   // k-blocks issued to the tensor core so far

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = p.C / 64;
  const int n_my = ((int)blockIdx.x < p.num_tiles) ? (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total = n_my * num_k;           // k-blocks this CTA walks through, over all of its tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_w);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); mbar_init(&a_empty[s], 1);
      mbar_init(&a_full[s], kMode == kIpNhwc ? 1 : 8);   // TMA bytes, or one arrival per converter warp of a group
    }
    for (int s = 0; s < 3; ++s) { mbar_init(&stg_full[s], 1); mbar_init(&stg_empty[s], 8); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
    *consumed = 0;
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 0) {
      // ===================== TMA producer: W k-blocks =====================
      if (lane == 0) {
        for (int g = 0; g < total; ++g) {
          const int sw = g % kWS, kb = g % num_k;
          mbar_wait(&w_empty[sw], ((g / kWS) & 1) ^ 1);
          mbar_expect_tx(&w_full[sw], kIpBBytes);
          tma_load_2d(smem_b + sw * kIpBBytes, &tma_w, &w_full[sw], kb * 64, 0);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
      int acc = 0, g = 0;
      uint32_t acc_phase = 0;
      for (int it = 0; it < n_my; ++it) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        for (int kb = 0; kb < num_k; ++kb, ++g) {
          const int sw = g % kWS, sa = g % kAS;
          mbar_wait(&w_full[sw], (g / kWS) & 1);
          mbar_wait(&a_full[sa], (g / kAS) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem_a + sa * kIpABytes));
            const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smem_b + sw * kIpBBytes));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&w_empty[sw]);
            umma_commit(&a_empty[sa]);
            if (kb == num_k - 1) umma_commit(&tfull_bar[acc]);
            *consumed = g + 1;
          }
          __syncwarp();
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    } else if (warp == 2) {
      if (lane == 0) {
        const size_t chunk = (size_t)64 * p.P;   // floats of one frame's k-block: 64 channels x P tokens, contiguous
        if constexpr (kMode == kIpNhwc) {
          // ===================== feature producer: the bf16 rows are the A operand — TMA tiles straight into the A ring =====================
          tma_prefetch_desc(&tma_a);
          for (int g = 0; g < total; ++g) {
            const int it = g / num_k, kb = g - it * num_k, sa = g % kAS;
            mbar_wait(&a_empty[sa], ((g / kAS) & 1) ^ 1);
            mbar_expect_tx(&a_full[sa], kIpABytes);
            tma_load_2d(smem_a + sa * kIpABytes, &tma_a, &a_full[sa], kb * 64, ((int)blockIdx.x + it * (int)gridDim.x) * 128);
          }
        } else if constexpr (kBulk) {
          // ===================== feature producer: bulk copies of the fp32 chunks into the staging ring =====================
          // (an additional L2 prefetch of the chunks ahead was measured and dropped: 469 us vs 449 us)
          for (int g = 0; g < total; ++g) {
            const int it = g / num_k, kb = g - it * num_k, s = g % kIpGroups;
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int f0 = tile * p.fpt;
            const int nf = p.F - f0 < p.fpt ? p.F - f0 : p.fpt;
            mbar_wait(&stg_empty[s], ((g / kIpGroups) & 1) ^ 1);
            mbar_expect_tx(&stg_full[s], (uint32_t)(nf * chunk * 4));
            for (int fl = 0; fl < nf; ++fl)
              bulk_copy_g2s(smem_stg + s * kIpStgBytes + fl * chunk * 4, p.in + ((size_t)(f0 + fl) * p.C + kb * 64) * p.P,
                            (uint32_t)(chunk * 4), &stg_full[s]);
          }
        } else if (p.l2_prefetch) {
          // ===================== L2 prefetcher (LDG form) =====================
          for (int g = 0; g < total; ++g) {
            while (g > *consumed + kIpPrefetchDist) __nanosleep(100);
            const int it = g / num_k, kb = g - it * num_k;
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            if (p.fpt > 0) {
              for (int fl = 0; fl < p.fpt; ++fl) {
                const int frame = tile * p.fpt + fl;
                if (frame < p.F) bulk_prefetch_l2(p.in + ((size_t)frame * p.C + kb * 64) * p.P, (uint32_t)(chunk * 4));
              }
            } else if (tile % p.tpf == 0) {        // the first slice of a frame fetches the chunk for all of its slices
              bulk_prefetch_l2(p.in + ((size_t)(tile / p.tpf) * p.C + kb * 64) * p.P, (uint32_t)(chunk * 4));
            }
          }
        }
      }
    }
  } else if (warp < 8) {
    // ===================== epilogue: TMEM row per thread → swizzled slab → coalesced row-segment stores =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    uint8_t* slab = smem_slab + quad * 4096;   // 32 rows x 32 fp32
    const int rl = lane >> 3, ul = lane & 7;   // read-back: row within a group of 4, 16-byte unit of the 128-byte slab row
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int it = 0; it < n_my; ++it) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      int frame, tok;
      const bool valid = ip_row(p, tile, r, frame, tok);
      const int my_orow = valid ? frame * p.S + p.tok0 + tok : -1;          // output row of this lane's TMEM row
      const int my_prow = !valid ? 0 : (p.pos_per_frame ? my_orow : p.tok0 + tok);
      int orow8[8], prow8[8];   // output / positional rows of the 8 slab rows this lane stores (rows itr*4 + rl)
#pragma unroll
      for (int itr = 0; itr < 8; ++itr) {
        orow8[itr] = __shfl_sync(0xffffffffu, my_orow, itr * 4 + rl);
        prow8[itr] = __shfl_sync(0xffffffffu, my_prow, itr * 4 + rl);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t raw[32];
        tmem_ld32(taddr + c * 32, raw);
        tmem_ld_wait();
        {
          uint8_t* rowp = slab + lane * 128;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<uint4*>(rowp + ((u ^ (lane & 7)) << 4)) = make_uint4(raw[4 * u], raw[4 * u + 1], raw[4 * u + 2], raw[4 * u + 3]);
        }
        __syncwarp();
        const int col = c * 32 + ul * 4;
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        // all positional loads of the chunk are issued before the first store (padding rows read row 0, unused)
        uint2 posv[8];
        if (p.XP != nullptr) {
#pragma unroll
          for (int itr = 0; itr < 8; ++itr)
            posv[itr] = __ldg(reinterpret_cast<const uint2*>(p.pos + (size_t)prow8[itr] * 256 + col));
        }
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {
          const int rr = itr * 4 + rl;
          if (orow8[itr] >= 0) {
            const float4 z = *reinterpret_cast<const float4*>(slab + rr * 128 + ((ul ^ (rr & 7)) << 4));
            const float y0 = z.x + b4.x, y1 = z.y + b4.y, y2 = z.z + b4.z, y3 = z.w + b4.w;
            const size_t o = (size_t)orow8[itr] * 256 + col;
            if (p.X32 != nullptr) __stcs(reinterpret_cast<float4*>(p.X32 + o), make_float4(y0, y1, y2, y3));
            __stcs(reinterpret_cast<uint2*>(p.X + o), make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3)));
            if (p.XP != nullptr) {
              const float2 pa = unpack_bf16(posv[itr].x), pb = unpack_bf16(posv[itr].y);
              __stcs(reinterpret_cast<uint2*>(p.XP + o), make_uint2(pack_bf16(y0 + pa.x, y1 + pa.y), pack_bf16(y2 + pb.x, y3 + pb.y)));
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if constexpr (kMode != kIpNhwc) {
    // ===================== converters: fp32 [c][p] (staging stage or global) → bf16 [p][c] swizzled (shared) =====================
    const int pt = threadIdx.x - 256;
    const int grp = pt >> 8;                  // this group owns the k-blocks with (running k-block index % kIpGroups) == grp
    const int r = pt & 127, half = (pt >> 7) & 1;
    const size_t cstride = (size_t)p.P;
    int it_cached = -1;
    const float* base = nullptr;              // LDG form: channel 0 of this thread's token in tile `it_cached`; nullptr = padding row
    int stg_off = -1;                         // bulk form: float offset of (frame-in-tile, channel 0, token) in a staging stage
    for (int g = grp; g < total; g += kIpGroups) {
      const int it = g / num_k, kb = g - it * num_k;
      if (it != it_cached) {
        int frame, tok;
        const bool valid = ip_row(p, (int)blockIdx.x + it * (int)gridDim.x, r, frame, tok);
        base = valid ? p.in + (size_t)frame * p.C * cstride + tok : nullptr;
        stg_off = valid ? (r / p.P) * 64 * p.P + tok : -1;
        it_cached = it;
      }
      float v[32];
      if constexpr (kBulk) {
        mbar_wait(&stg_full[grp], ((g / kIpGroups) & 1));
        if (stg_off >= 0) {
          const float* src = reinterpret_cast<const float*>(smem_stg + grp * kIpStgBytes) + stg_off + half * 32 * p.P;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = src[i * p.P];
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&stg_empty[grp]);   // the stage can be refilled while this k-block is converted
      } else {
        if (base != nullptr) {
          const float* src = base + (size_t)(kb * 64 + half * 32) * cstride;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __ldg(src + (size_t)i * cstride);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
      }
      const int sa = g % kAS;
      mbar_wait(&a_empty[sa], ((g / kAS) & 1) ^ 1);
      uint8_t* rowp = smem_a + sa * kIpABytes + r * 128;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        *reinterpret_cast<uint4*>(rowp + (((half * 4 + u) ^ (r & 7)) << 4)) =
            make_uint4(pack_bf16(v[8 * u], v[8 * u + 1]), pack_bf16(v[8 * u + 2], v[8 * u + 3]),
                       pack_bf16(v[8 * u + 4], v[8 * u + 5]), pack_bf16(v[8 * u + 6], v[8 * u + 7]));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[sa]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------ host
CUtensorMap make_tmap_2d(const void* ptr, int rows, int cols, int ld, int box_rows, bool f32);  // gemm_tc.cu
int device_sm_count();
void count_gemm_launch();

bool input_proj_supported(int C) { return C >= 64 && C % 64 == 0; }

template <int kMode>
static void launch_ip(const CUtensorMap& tw, const CUtensorMap& ta, const IpParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    VG_CUDA(cudaFuncSetAttribute(input_proj_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, IpCfg<kMode>::kSmem));
    attr_set = true;
  }
  const int sms = device_sm_count();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  input_proj_kernel<kMode><<<grid, IpCfg<kMode>::kThreads, IpCfg<kMode>::kSmem, stream>>>(tw, ta, p);
  VG_CUDA(cudaGetLastError());
}

// X / X32 / XP rows [f*S + tok0, +P) of every frame f < F  <-  W in[f] + b  (+ pos); see the header comment.
void input_proj(const float* in, int C, const bf16* W, const float* bias, const bf16* pos, int pos_frames, bf16* X, float* X32,
                bf16* XP, int F, int S, int tok0, int P, cudaStream_t stream) {
  VG_CHECK(input_proj_supported(C), "input_proj: the channel count must be a multiple of 64");
  VG_CHECK(in && W && bias && X && F > 0 && P > 0 && tok0 >= 0 && tok0 + P <= S, "input_proj: bad arguments");
  VG_CHECK(XP == nullptr || pos != nullptr, "input_proj: XP needs the positional rows");
  VG_CHECK(pos_frames == 1 || pos_frames == F, "input_proj: pos_frames must be 1 or F");
  IpParams p;
  p.in = in; p.bias = bias; p.pos = pos; p.X = X; p.X32 = X32; p.XP = XP;
  p.F = F; p.C = C; p.P = P; p.S = S; p.tok0 = tok0; p.pos_per_frame = pos_frames > 1 ? 1 : 0;
  if (P <= 128) { p.fpt = 128 / P; p.tpf = 1; p.num_tiles = (F + p.fpt - 1) / p.fpt; }
  else { p.fpt = 0; p.tpf = (P + 127) / 128; p.num_tiles = F * p.tpf; }
  const bool aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  p.l2_prefetch = aligned ? 1 : 0;
  if (const char* e = getenv("VGQA_IP_PREFETCH")) p.l2_prefetch = p.l2_prefetch && e[0] != '0';
  // whole frames per tile + aligned input → the TMA-engine (bulk copy) form; VGQA_IP_BULK=0 forces the LDG form (A/B runs)
  bool bulk = aligned && P <= 128;
  if (const char* e = getenv("VGQA_IP_BULK")) bulk = bulk && e[0] != '0';
  CUtensorMap tw = make_tmap_2d(W, 256, C, C, 256, false);
  if (bulk) launch_ip<kIpBulk>(tw, tw, p, stream);
  else launch_ip<kIpLdg>(tw, tw, p, stream);
  count_gemm_launch();
}

// The same projection from a channels-last bf16 map  in[f, p, c]  ([F*P, C] rows — what a bf16 channels_last backbone emits):
// the rows are the K-major A operand, so they go from HBM into the UMMA tile by TMA with no conversion at half the bytes.
void input_proj_nhwc(const bf16* in, int C, const bf16* W, const float* bias, const bf16* pos, int pos_frames, bf16* X, float* X32,
                     bf16* XP, int F, int S, int tok0, int P, cudaStream_t stream) {
  VG_CHECK(input_proj_supported(C), "input_proj: the channel count must be a multiple of 64");
  VG_CHECK(in && W && bias && X && F > 0 && P > 0 && tok0 >= 0 && tok0 + P <= S, "input_proj: bad arguments");
  VG_CHECK(XP == nullptr || pos != nullptr, "input_proj: XP needs the positional rows");
  VG_CHECK(pos_frames == 1 || pos_frames == F, "input_proj: pos_frames must be 1 or F");
  VG_CHECK((long long)F * P < (1ll << 31), "input_proj: too many rows");
  IpParams p;
  p.in = nullptr; p.bias = bias; p.pos = pos; p.X = X; p.X32 = X32; p.XP = XP;
  p.F = F; p.C = C; p.P = P; p.S = S; p.tok0 = tok0; p.pos_per_frame = pos_frames > 1 ? 1 : 0;
  p.fpt = 0; p.tpf = 0; p.num_tiles = (F * P + 127) / 128; p.l2_prefetch = 0;
  CUtensorMap tw = make_tmap_2d(W, 256, C, C, 256, false);
  CUtensorMap ta = make_tmap_2d(in, F * P, C, C, 128, false);
  launch_ip<kIpNhwc>(tw, ta, p, stream);
  count_gemm_launch();
}

// ------------------------------------------------------------------------------------------ fp32 → bf16 rows (text_raw)
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}
void f32_to_bf16(const float* in, bf16* out, size_t n, cudaStream_t st) {
  VG_CHECK(n % 4 == 0, "f32_to_bf16: length must be a multiple of 4");
  const size_t n4 = n / 4;
  f32_to_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(in, out, n4);
  VG_CUDA(cudaGetLastError());
}

}  // namespace vg
