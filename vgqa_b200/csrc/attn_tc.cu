// Per-frame encoder self-attention on tcgen05 / TMEM / TMA (S <= 128 tokens per frame, 8 heads, head_dim 32).
//
// Reference: F.multi_head_attention_forward inside TransformerEncoderLayer (vgqa/core/decoder/modal_encoder.py:172).
// Input is the packed in-projection output QKV[R, 768] (q | k | v), output AO[R, 256] (heads concatenated).
//
// One persistent CTA per SM walks over frames; the unit of work is (frame, head):
//   warp 0      TMA producer: Q, K, V head slices (128 rows x 64 B, 64B-swizzled) through a 3D tensor map
//               (cols, tokens, frames) whose token dimension is the true S, so rows >= S are zero-filled on load and
//               clipped on store — no cross-frame contamination, no tail code.
//   warp 1      tcgen05.mma issuer:  S = Q K^T (128x128x32, 2 UMMAs, K-major SW64 operands) into TMEM, and
//               O = P V (128x32x128, 8 UMMAs; P is a K-major SW128 smem operand written by the softmax warps,
//               V is an MN-major SW64 operand straight from the TMA tile).
//   warps 2..17 FOUR softmax warpgroups taking units round-robin, one score row per thread (TMEM lane):
//               row max / exp2 without any shuffle, P packed to bf16 in shared memory, then the O epilogue
//               (1/sum scaling, bf16, TMA store).
// The softmax of one unit is a ~4.5 k-cycle dependent chain (in-kernel timeline: max pass 1.25 k, exp pass 3.3 k) that neither
// MUFU nor issue bandwidth explains, so throughput comes from units in flight: four S tiles live in TMEM (4 x 128 columns;
// the O tile of a unit is written over the first 48 columns of its own S tile once the softmax has consumed it), Q|K and V
// travel through separate rings (V must outlive the softmax), and the MMA warp issues P·V three units behind Q·K^T.
#include "attn_tc.cuh"

namespace vg {

__global__ void __launch_bounds__(kAtThreads, 1)
enc_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_o,
                   const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_qk = smem;                                  // [kAtQkStages][Q|K]
  uint8_t* s_v = s_qk + kAtQkStages * kAtQkBytes;        // [kAtVStages]
  uint8_t* s_p = s_v + kAtVStages * kAtVBytes;           // [kAtWgs][2 atoms]
  uint8_t* s_ones = s_p + kAtWgs * kAtPBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + kAtOnesBytes);
  uint64_t* qk_full = bars;                       // [kAtQkStages]
  uint64_t* qk_empty = qk_full + kAtQkStages;     // [kAtQkStages]
  uint64_t* v_full = qk_empty + kAtQkStages;      // [kAtVStages]
  uint64_t* v_empty = v_full + kAtVStages;        // [kAtVStages]
  uint64_t* s_full = v_empty + kAtVStages;        // [kAtWgs]
  uint64_t* p_full = s_full + kAtWgs;             // [kAtWgs]
  uint64_t* o_full = p_full + kAtWgs;             // [kAtWgs]
  uint64_t* t_free = o_full + kAtWgs;             // [kAtWgs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_free + kAtWgs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  int my_frames = 0;
  for (int f = blockIdx.x; f < p.F; f += gridDim.x) ++my_frames;
  const int num_units = my_frames * 8;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_o);
    for (int b = 0; b < kAtQkStages; ++b) { mbar_init(&qk_full[b], 1); mbar_init(&qk_empty[b], 1); }
    for (int b = 0; b < kAtVStages; ++b) { mbar_init(&v_full[b], 1); mbar_init(&v_empty[b], 1); }
    for (int b = 0; b < kAtWgs; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 128);
      mbar_init(&o_full[b], 1);
      mbar_init(&t_free[b], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (warp >= 2) {  // ones tile (bf16 1.0 = 0x3F80); uniform, so the swizzle pattern is irrelevant
    for (int i = threadIdx.x - 64; i < kAtOnesBytes / 16; i += kAtWgs * 128)
      reinterpret_cast<uint4*>(s_ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int u = 0; u < num_units; ++u) {
        const int sq = u % kAtQkStages, sv = u % kAtVStages;
        const int f = blockIdx.x + (u >> 3) * gridDim.x, h = u & 7;
        mbar_wait(&qk_empty[sq], ((u / kAtQkStages) & 1) ^ 1);
        mbar_expect_tx(&qk_full[sq], kAtQkBytes);
        uint8_t* dst = s_qk + sq * kAtQkBytes;
        tma_load_3d(dst, &tm_qkv, &qk_full[sq], h * 32, 0, f);
        tma_load_3d(dst + 8192, &tm_qkv, &qk_full[sq], 256 + h * 32, 0, f);
        mbar_wait(&v_empty[sv], ((u / kAtVStages) & 1) ^ 1);
        mbar_expect_tx(&v_full[sv], kAtVBytes);
        tma_load_3d(s_v + sv * kAtVBytes, &tm_qkv, &v_full[sv], 512 + h * 32, 0, f);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 48) | (1u << 16);  // B = [V | ones] is MN-major, N = 32 + 16
    for (int u = 0; u < num_units + kAtLag; ++u) {
      if (u < num_units) {
        const int b = u % kAtWgs, sq = u % kAtQkStages;
        mbar_wait(&qk_full[sq], (u / kAtQkStages) & 1);
        mbar_wait(&t_free[b], ((u / kAtWgs) & 1) ^ 1);   // the O tile of unit u - kAtWgs (aliased on this S tile) has been read
        tc_fence_after();
        if (elect_one()) {
          const uint32_t q_addr = smem_u32(s_qk + sq * kAtQkBytes);
          const uint64_t qd = umma_desc_sw64_kmajor(q_addr), kd = umma_desc_sw64_kmajor(q_addr + 8192);
          umma_bf16(tmem_base + b * 128, qd, kd, idesc_s, 0u);
          umma_bf16(tmem_base + b * 128, qd + 2, kd + 2, idesc_s, 1u);
          umma_commit(&s_full[b]);
          umma_commit(&qk_empty[sq]);
        }
        __syncwarp();
      }
      if (u >= kAtLag) {
        const int v = u - kAtLag, b = v % kAtWgs, sv = v % kAtVStages;
        mbar_wait(&v_full[sv], (v / kAtVStages) & 1);
        mbar_wait(&p_full[b], (v / kAtWgs) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t v_addr = smem_u32(s_v + sv * kAtVBytes);
          const uint32_t p_addr = smem_u32(s_p + b * kAtPBytes);
          const uint64_t vd = umma_desc_sw64_mnmajor(v_addr, smem_u32(s_ones) - v_addr);
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // 8 x 16 keys; O lands on the first 48 columns of the (consumed) S tile
            const uint64_t pd = umma_desc_sw128_kmajor(p_addr + (k >> 2) * 16384) + 2 * (k & 3);
            umma_bf16(tmem_base + b * 128, pd, vd + (uint64_t)((k * 1024) >> 4), idesc_o, k ? 1u : 0u);
          }
          umma_commit(&o_full[b]);
          umma_commit(&v_empty[sv]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax + output warpgroups =====================
    const int wg = (warp - 2) >> 2;          // handles units u with u % kAtWgs == wg
    const int quad = warp & 3;
    const int row = quad * 32 + lane;        // score row / token index within the frame
    const int wg_tid = (warp - 2 - wg * 4) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    uint8_t* p_buf = s_p + wg * kAtPBytes;
    uint8_t* o_buf = p_buf;                  // O staging reuses the head of this warpgroup's P tile (dead once P·V is done)
    for (int u = wg; u < num_units; u += kAtWgs) {
      const int f = blockIdx.x + (u >> 3) * gridDim.x, h = u & 7;
      const uint32_t ph = (u / kAtWgs) & 1;
      const uint8_t* km = p.kmask ? p.kmask + (size_t)f * S : nullptr;
      mbar_wait(&s_full[wg], ph);
      tc_fence_after();
      const uint32_t t_s = tmem_base + lane_base + wg * 128;
      uint32_t raw[32];
      // Dead keys (index >= S, or padded by the key mask) are set to -inf once per chunk; everything after that is
      // branch-free: max, then p = ex2(s*scale - base) with ex2(-inf) = 0.
      auto load_chunk = [&](int c, float (&x)[32]) {
        tmem_ld32(t_s + c * 32, raw);
        tmem_ld_wait();
        if (km != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int key = c * 32 + i;
            x[i] = (key >= S || km[key] != 0) ? -INFINITY : __uint_as_float(raw[i]);
          }
        } else if (c * 32 + 32 > S) {
          const int nvalid = S - c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = i < nvalid ? __uint_as_float(raw[i]) : -INFINITY;
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(raw[i]);
        }
      };
      float x[32];
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c * 32 < S; ++c) {
        load_chunk(c, x);
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, x[i]);
      }
      const float nbase = -(mx == -INFINITY ? 0.f : mx) * p.scale_log2e;
      const uint32_t sw = static_cast<uint32_t>(row & 7) << 4;
      if (wg_tid == 0) tma_store_wait_read<0>();  // the previous unit's output store has drained the head of p_buf
      named_bar_sync(1 + wg, 128);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
        if (c * 32 < S) {
          load_chunk(c, x);
#pragma unroll
          for (int i = 0; i < 16; ++i)   // two exponentials per MUFU op, result already packed bf16x2
            pk[i] = ex2_bf16x2(pack_bf16(fmaf(x[2 * i], p.scale_log2e, nbase), fmaf(x[2 * i + 1], p.scale_log2e, nbase)));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
        }
        uint8_t* rowp = p_buf + (c >> 1) * 16384 + row * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(rowp + ((static_cast<uint32_t>((c & 1) * 4 + i) << 4) ^ sw)) =
              make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[wg]);
      // O epilogue
      mbar_wait(&o_full[wg], ph);
      tc_fence_after();
      tmem_ld32(tmem_base + lane_base + wg * 128, raw);
      const float sum = __uint_as_float(tmem_ld1(tmem_base + lane_base + wg * 128 + 32));  // P · ones
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_free[wg]);   // this S/O tile may be overwritten by unit u + kAtWgs
      const float inv = 1.f / sum;
      {
        uint8_t* rowp = o_buf + row * 64;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(rowp + ((i ^ ((row >> 1) & 3)) << 4)) =
              make_uint4(pack_bf16(__uint_as_float(raw[8 * i]) * inv, __uint_as_float(raw[8 * i + 1]) * inv),
                         pack_bf16(__uint_as_float(raw[8 * i + 2]) * inv, __uint_as_float(raw[8 * i + 3]) * inv),
                         pack_bf16(__uint_as_float(raw[8 * i + 4]) * inv, __uint_as_float(raw[8 * i + 5]) * inv),
                         pack_bf16(__uint_as_float(raw[8 * i + 6]) * inv, __uint_as_float(raw[8 * i + 7]) * inv));
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + wg, 128);
      if (wg_tid == 0) {
        tma_store_3d(&tm_o, o_buf, h * 32, 0, f);
        tma_store_commit();
      }
    }
    if (wg_tid == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
void* tensor_map_encode_fn();  // gemm_tc.cu
int device_sm_count();

// [F frames][S tokens][cols] bf16 with row stride ld; box = 32 cols x 128 tokens x 1 frame, 64B swizzle.
CUtensorMap make_tmap_frames(const bf16* ptr, int F, int S, int cols, int ld) {
  CUtensorMap m;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)S, (cuuint64_t)F};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)S * ld * 2};
  cuuint32_t box[3] = {32, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<PFN_encodeTiled_t>(tensor_map_encode_fn())(
      &m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(ptr), gdim, gstr, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3D) failed with CUresult " + std::to_string((int)r));
  return m;
}

void enc_attn_tc_long(const bf16* QKV, bf16* AO, int F, int S, const uint8_t* kmask, float scale, cudaStream_t stream);  // attn_tc_long.cu

bool enc_attn_tc_supported(int S) { return S >= 1 && S <= 4096; }

// AO[f*S + s, h*32 + d] = softmax_s'(scale * q·k) v  over the S tokens of frame f; QKV is [F*S, 768] (q | k | v).
void enc_attn_tc(const bf16* QKV, bf16* AO, int F, int S, const uint8_t* kmask, float scale, cudaStream_t stream) {
  VG_CHECK(enc_attn_tc_supported(S) && F > 0, "enc_attn_tc: S must be in [1,4096]");
  if (S > 128) {   // several 128-key tiles per frame: online-softmax variant
    enc_attn_tc_long(QKV, AO, F, S, kmask, scale, stream);
    return;
  }
  static bool attr_set = false;
  if (!attr_set) {
    VG_CUDA(cudaFuncSetAttribute(enc_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    attr_set = true;
  }
  CUtensorMap tq = make_tmap_frames(QKV, F, S, 768, 768);
  CUtensorMap to = make_tmap_frames(AO, F, S, 256, 256);
  AttnTcParams p;
  p.kmask = kmask; p.S = S; p.F = F; p.scale_log2e = scale * 1.4426950408889634f;
  const int grid = F < device_sm_count() ? F : device_sm_count();
  enc_attn_tc_kernel<<<grid, kAtThreads, kAtSmem, stream>>>(tq, to, p);
  VG_CUDA(cudaGetLastError());
}

}  // namespace vg
