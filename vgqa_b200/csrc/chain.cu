// Row-tile GEMM CHAINS of the decoders on tcgen05: one launch runs a whole sequence of dependent Linear layers over a tile of 128
// query rows ((clip, frame) rows of TimeDecoder / PosDecoder, query_decoder.py:208-486), with the activations passed from one
// GEMM to the next through shared memory instead of through one kernel launch + HBM round trip per Linear:
//
//   chain "tail" of a decoder layer   ctx8 → [vo + LN3] → x2 → [linear1 + ReLU → linear2 + LN4] → tgt
//        TimeDecoder:  → time_decoder.norm → hs[l];  → next layer's packed in-projection (+ time table)          (:472-485,:412,:466)
//        PosDecoder:   → bbox_embed (256 → 256 → 256 → 4) → sigmoid → reference boxes of the next layer           (:368-374,:188-192)
//
// Every query row only ever meets its own values in these ops (the row-coupling ops — temporal self-attention, cross-attention —
// stay separate launches that spread over the whole GPU), so a CTA owns its 128 rows for the whole chain and nothing is exchanged
// between CTAs.  What bounds a chain is the stream of WEIGHTS through the SM (1 – 3.4 MB per chain): they are pre-tiled at pack
// time into 16 KB images of the 128B-swizzled K-major UMMA operand layout ([128 N-rows x 64 K] bf16), so the producer warp
// fetches one tile with ONE bulk copy (cp.async.bulk, no tensor map) and runs ahead of the tensor core through a 5-slot ring —
// across the GEMMs of the chain, because the weight addresses do not depend on data.
//
//   warp 0      producer: weight tiles (bulk copies), streamed A tiles of the K = 2048 input (TMA), initial activation loads
//   warp 1      tcgen05.mma issuer: D[128 x 128] += A[128 x 64] · W[128 x 64]^T per tile, two 256-column TMEM accumulators
//   warps 2..5  epilogue, one accumulator row per thread: bias / table / activation / residual + LayerNorm (two-pass from TMEM,
//               the normalised row parked in TMEM between the passes) / second LayerNorm / 4-wide head, and the bf16 rows
//               of the next GEMM's A operand written straight into the swizzled shared-memory tiles (ACT buffers)
//
// The FFN pair follows ffn_fused.cu's scheme on one CTA: the hidden activation is walked in 128-wide chunks, H_c = relu(x W1_c^T)
// accumulates in one half of the second accumulator, is converted to bf16 into one half of ACT1 and immediately consumed as the
// K-slice of Out += H_c · W2_c^T; the first GEMM runs two chunks ahead of the second.
#include <math.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "chain.h"
#include "ptx.cuh"

namespace vg {

static constexpr int CH_TILE = 16384;   // bytes of one operand tile: 128 rows x 64 bf16, rows of 128 B, 128B swizzle
static constexpr int CH_SLOTS = 5;      // ring of weight / streamed-A tiles
static constexpr int CH_HEADW = 4 * 256 * 4;
static constexpr int CH_SMEM = 8 * CH_TILE + CH_SLOTS * CH_TILE + CH_HEADW + 512 + 1024;

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float ch_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// ------------------------------------------------------------------------------------------------ epilogue helpers
struct EpiCtx {
  const float* sine_inv;   // 64 frequencies of the box sine embedding (kernel parameter space)
  uint32_t tmem_row;   // TMEM address of this thread's lane, column 0
  uint8_t* act;        // ACT tiles (8 x 16 KB)
  const float* headw;  // staged head weights [head_n][256]
  int grow;            // global row
  int lrow;            // row inside the tile
  bool valid;
  int t;               // frame index of the row (table row)
};

// bf16 row values of 32 columns [col0, col0+32) → the swizzled A-operand tile that holds them
__device__ __forceinline__ void store_act32(const EpiCtx& e, int blk0, int col0, const float (&v)[32]) {
  uint8_t* rowp = e.act + (size_t)(blk0 + (col0 >> 6)) * CH_TILE + e.lrow * 128;
  const int c0 = (col0 & 63) >> 3, swz = e.lrow & 7;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<uint4*>(rowp + (((c0 + i) ^ swz) << 4)) =
        make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]), pack_bf16(v[8 * i + 4], v[8 * i + 5]),
                   pack_bf16(v[8 * i + 6], v[8 * i + 7]));
}
__device__ __forceinline__ void store_bf16_32(bf16* dst, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    reinterpret_cast<uint4*>(dst)[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                                  pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
}
__device__ __forceinline__ void store_f32_32(float* dst, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void add_vec32(float (&v)[32], const float* src) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(src) + i);
    v[4 * i] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
  }
}

// Plain epilogue of one accumulator chunk (ncols = 128 or 256 columns starting at column `ncol0` of the op's output):
// v = act(acc + bias + table[t]); outputs: ACT tiles / global bf16 / global fp32 / 4-wide head.
__device__ void epi_plain(const ChOp& op, const EpiCtx& e, uint32_t tacc, int ncol0, int ncols) {
  uint32_t raw[32];
  float v[32];
  float hacc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = 0; c < ncols / 32; ++c) {
    tmem_ld32(tacc + c * 32, raw);
    tmem_ld_wait();
    const int col = ncol0 + c * 32;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
    if (op.bias != nullptr) add_vec32(v, op.bias + col);
    if (op.table != nullptr) add_vec32(v, op.table + (size_t)e.t * op.table_ld + col);
    if (op.act == ACT_RELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (op.act == ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ch_gelu(v[i]);
    }
    if (op.out_act >= 0) store_act32(e, op.out_act, col, v);
    if (e.valid) {
      if (op.out_bf16 != nullptr) store_bf16_32(op.out_bf16 + (size_t)e.grow * op.ld_out + col, v);
      if (op.out_f32 != nullptr) store_f32_32(op.out_f32 + (size_t)e.grow * op.ld_out_f32 + col, v);
    }
    if (op.head_n > 0) {   // y[j] = sum_c bf16(v[c]) * w[j][c]  (the unfused path feeds the head with bf16 rows)
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __bfloat162float(__float2bfloat16(v[i]));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < op.head_n) {
          const float* w = e.headw + j * 256 + col;
          float a = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = *reinterpret_cast<const float4*>(w + 4 * i);
            a = fmaf(v[4 * i], q.x, a); a = fmaf(v[4 * i + 1], q.y, a); a = fmaf(v[4 * i + 2], q.z, a); a = fmaf(v[4 * i + 3], q.w, a);
          }
          hacc[j] += a;
        }
      }
    }
  }
  if (op.head_n > 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < op.head_n) {
        float y = hacc[j] + __ldg(op.head_b + j);
        if (op.head_act == 1) y = 1.f / (1.f + expf(-y));
        hacc[j] = y;
        if (e.valid) op.head_out[(size_t)e.grow * op.head_n + j] = y;
      }
    }
    if (op.gen_sine) {   // boxes (cx, cy, w, h) → 4 x 128 sine features in the order (y, x, w, h): sin on even, cos on odd indices
#pragma unroll 1
      for (int grp = 0; grp < 4; ++grp) {
        const float x = grp == 0 ? hacc[1] : (grp == 1 ? hacc[0] : (grp == 2 ? hacc[2] : hacc[3]));
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float sv, cv;
            __sincosf(x * e.sine_inv[c * 16 + i], &sv, &cv);
            v[2 * i] = sv; v[2 * i + 1] = cv;
          }
          store_act32(e, 0, grp * 128 + c * 32, v);
          if (e.valid && op.sine_out != nullptr) store_bf16_32(op.sine_out + (size_t)e.grow * 512 + grp * 128 + c * 32, v);
        }
      }
    }
  }
}

// Residual + LayerNorm epilogue of a 256-column accumulator: v = acc + bias + res32[row]; y = LN(v) → out32 / ACT tiles / global
// bf16; optional second LayerNorm z = LN2(y) → out2 (global bf16).  The row is parked in TMEM between the passes.
__device__ void epi_ln(const ChOp& op, const EpiCtx& e, uint32_t tacc, const float* bias) {
  uint32_t raw[32];
  float v[32];
  float shift = 0.f, s1 = 0.f, s2 = 0.f;
  for (int c = 0; c < 8; ++c) {
    tmem_ld32(tacc + c * 32, raw);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
    if (bias != nullptr) add_vec32(v, bias + c * 32);
    if (op.act == ACT_RELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (op.act == ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ch_gelu(v[i]);
    }
    if (op.res32 != nullptr && e.valid) add_vec32(v, op.res32 + (size_t)e.grow * 256 + c * 32);
    if (c == 0) shift = v[0];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float d = v[i] - shift;
      s1 += d;
      s2 = fmaf(d, d, s2);
      raw[i] = __float_as_uint(v[i]);
    }
    tmem_st32(tacc + c * 32, raw);
  }
  tmem_st_wait();
  float dm = s1 * (1.0f / 256.f);
  float mean = shift + dm;
  float rstd = rsqrtf(fmaxf(s2 * (1.0f / 256.f) - dm * dm, 0.f) + op.ln_eps);
  const bool second = op.ln2_w != nullptr;
  shift = 0.f; s1 = 0.f; s2 = 0.f;
  for (int c = 0; c < 8; ++c) {
    tmem_ld32(tacc + c * 32, raw);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(op.ln_w + c * 32) + i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(op.ln_b + c * 32) + i);
      v[4 * i] = (__uint_as_float(raw[4 * i]) - mean) * rstd * w.x + b.x;
      v[4 * i + 1] = (__uint_as_float(raw[4 * i + 1]) - mean) * rstd * w.y + b.y;
      v[4 * i + 2] = (__uint_as_float(raw[4 * i + 2]) - mean) * rstd * w.z + b.z;
      v[4 * i + 3] = (__uint_as_float(raw[4 * i + 3]) - mean) * rstd * w.w + b.w;
    }
    if (op.out_act >= 0) store_act32(e, op.out_act, c * 32, v);
    if (e.valid) {
      if (op.out32 != nullptr) store_f32_32(op.out32 + (size_t)e.grow * 256 + c * 32, v);
      if (op.out_bf16 != nullptr) store_bf16_32(op.out_bf16 + (size_t)e.grow * op.ld_out + c * 32, v);
    }
    if (second) {
      if (c == 0) shift = v[0];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float d = v[i] - shift;
        s1 += d;
        s2 = fmaf(d, d, s2);
        raw[i] = __float_as_uint(v[i]);
      }
      tmem_st32(tacc + c * 32, raw);
    }
  }
  if (second) {
    tmem_st_wait();
    dm = s1 * (1.0f / 256.f);
    mean = shift + dm;
    rstd = rsqrtf(fmaxf(s2 * (1.0f / 256.f) - dm * dm, 0.f) + op.ln2_eps);
    for (int c = 0; c < 8; ++c) {
      tmem_ld32(tacc + c * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(op.ln2_w + c * 32) + i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(op.ln2_b + c * 32) + i);
        v[4 * i] = (__uint_as_float(raw[4 * i]) - mean) * rstd * w.x + b.x;
        v[4 * i + 1] = (__uint_as_float(raw[4 * i + 1]) - mean) * rstd * w.y + b.y;
        v[4 * i + 2] = (__uint_as_float(raw[4 * i + 2]) - mean) * rstd * w.z + b.z;
        v[4 * i + 3] = (__uint_as_float(raw[4 * i + 3]) - mean) * rstd * w.w + b.w;
      }
      if (e.valid) store_bf16_32(op.out2 + (size_t)e.grow * op.ld_out2 + c * 32, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(192, 1) chain_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act = smem;                          // ACT0 = tiles 0..3, ACT1 = tiles 4..7
  uint8_t* ring = smem + 8 * CH_TILE;
  float* headw = reinterpret_cast<float*>(ring + CH_SLOTS * CH_TILE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(headw) + CH_HEADW);
  uint64_t* full = bars;                        // [CH_SLOTS] tile landed
  uint64_t* empty = full + CH_SLOTS;            // [CH_SLOTS] the MMAs reading the tile are done
  uint64_t* acc_full = empty + CH_SLOTS;        // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2] (4 epilogue warps)
  uint64_t* act_tma = acc_empty + 2;            // [2] an ACT buffer was loaded by TMA
  uint64_t* act_epi = act_tma + 2;              // [2] an ACT buffer was written by the epilogue (4 warps)
  uint64_t* hacc_full = act_epi + 2;            // [2] FFN: hidden chunk accumulated
  uint64_t* hacc_empty = hacc_full + 2;         // [2] FFN: hidden accumulator half drained (4 warps)
  uint64_t* h_full = hacc_empty + 2;            // [2] FFN: bf16 hidden chunk written to ACT1 half (4 warps)
  uint64_t* h_free = h_full + 2;                // [2] FFN: second-GEMM MMAs reading the ACT1 half are done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int row0 = tile * 128;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < P.n_tm; ++i) tma_prefetch_desc(&P.tm[i]);
    for (int s = 0; s < CH_SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); mbar_init(&act_tma[a], 1); mbar_init(&act_epi[a], 4);
      mbar_init(&hacc_full[a], 1); mbar_init(&hacc_empty[a], 4); mbar_init(&h_full[a], 4); mbar_init(&h_free[a], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (warp >= 2) {   // head weights (constant): staged once
    for (int oi = 0; oi < P.n_ops; ++oi)
      if (P.ops[oi].head_n > 0)
        for (int i = threadIdx.x - 64; i < P.ops[oi].head_n * 256; i += 128) headw[i] = __ldg(P.ops[oi].head_w + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== producer
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      auto next = [&]() { if (++slot == CH_SLOTS) { slot = 0; ph ^= 1; } };
      auto push_w = [&](const bf16* tile_src) {
        mbar_wait(&empty[slot], ph ^ 1);
        mbar_expect_tx(&full[slot], CH_TILE);
        bulk_copy_g2s(ring + slot * CH_TILE, tile_src, CH_TILE, &full[slot]);
        next();
      };
      auto push_a = [&](const CUtensorMap* tm, int col) {
        mbar_wait(&empty[slot], ph ^ 1);
        mbar_expect_tx(&full[slot], CH_TILE);
        tma_load_2d(ring + slot * CH_TILE, tm, &full[slot], col, row0);
        next();
      };
      bool waited = false;
      auto dep_wait = [&]() {   // activations of earlier kernels: only behind the PDL wait; weights may be fetched before it
        if (!waited) { pdl_wait(); pdl_launch_dependents(); waited = true; }
      };
      for (int oi = 0; oi < P.n_ops; ++oi) {
        const ChOp& op = P.ops[oi];
        if (op.kind == CH_LOAD) {
          dep_wait();
          const int buf = op.a_blk >> 2;
          mbar_expect_tx(&act_tma[buf], op.nkb * CH_TILE);
          for (int kb = 0; kb < op.nkb; ++kb)
            tma_load_2d(act + (size_t)(op.a_blk + kb) * CH_TILE, &P.tm[op.a_tm], &act_tma[buf], op.a_col0 + kb * 64, row0);
        } else if (op.kind == CH_GEMM) {
          const int chunks = op.nblk / op.cn;
          for (int ch = 0; ch < chunks; ++ch)
            for (int kb = 0; kb < op.nkb; ++kb) {
              if (op.a_kind == CH_A_STREAM) { dep_wait(); push_a(&P.tm[op.a_tm], op.a_col0 + kb * 64); }
              for (int j = 0; j < op.cn; ++j) push_w(op.w + ((size_t)(ch * op.cn + j) * op.nkb + kb) * (CH_TILE / 2));
            }
        } else {   // CH_FFN: H(0), H(1), then ff2(c), H(c+2) ...
          const int nc = op.ff_chunks, nkb2 = nc * 2;
          auto push_h = [&](int c) { for (int kb = 0; kb < 4; ++kb) push_w(op.w + ((size_t)c * 4 + kb) * (CH_TILE / 2)); };
          push_h(0);
          if (nc > 1) push_h(1);
          for (int c = 0; c < nc; ++c) {
            for (int kbl = 0; kbl < 2; ++kbl)
              for (int nb = 0; nb < 2; ++nb) push_w(op.w2 + ((size_t)nb * nkb2 + 2 * c + kbl) * (CH_TILE / 2));
            if (c + 2 < nc) push_h(c + 2);
          }
        }
      }
      dep_wait();
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
    int slot = 0;
    uint32_t ph = 0;
    uint32_t acc_use[2] = {0, 0}, tma_use[2] = {0, 0}, epi_use[2] = {0, 0};
    uint32_t hbase = 0;   // hidden chunks of earlier FFN ops
    const uint32_t act_addr = smem_u32(act), ring_addr = smem_u32(ring);
    // one [128 x 64] x [128 x 64]^T tile product: 4 MMAs of K = 16
    auto mma_tile = [&](uint32_t a_smem, uint32_t w_smem, uint32_t d_tmem, bool first) {
      const uint64_t ad = umma_desc_sw128_kmajor(a_smem), bd = umma_desc_sw128_kmajor(w_smem);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
    };
    auto wait_a = [&](const ChOp& op) {
      const int buf = op.a_blk >> 2;
      if (op.a_wait == CH_WAIT_TMA) { mbar_wait(&act_tma[buf], tma_use[buf] & 1); ++tma_use[buf]; }
      else if (op.a_wait == CH_WAIT_EPI) { mbar_wait(&act_epi[buf], epi_use[buf] & 1); ++epi_use[buf]; }
      tc_fence_after();
    };
    for (int oi = 0; oi < P.n_ops; ++oi) {
      const ChOp& op = P.ops[oi];
      if (op.kind == CH_LOAD) continue;
      if (op.kind == CH_GEMM) {
        const int chunks = op.nblk / op.cn;
        int acc = op.acc0;
        for (int ch = 0; ch < chunks; ++ch, acc ^= 1) {
          mbar_wait(&acc_empty[acc], (acc_use[acc] & 1) ^ 1);
          ++acc_use[acc];
          if (ch == 0) wait_a(op);
          tc_fence_after();
          for (int kb = 0; kb < op.nkb; ++kb) {
            uint32_t a_smem;
            int a_slot = -1;
            if (op.a_kind == CH_A_STREAM) {
              mbar_wait(&full[slot], ph);
              a_slot = slot;
              a_smem = ring_addr + slot * CH_TILE;
              if (++slot == CH_SLOTS) { slot = 0; ph ^= 1; }
            } else {
              a_smem = act_addr + (op.a_blk + kb) * CH_TILE;
            }
            for (int j = 0; j < op.cn; ++j) {
              mbar_wait(&full[slot], ph);
              tc_fence_after();
              if (elect_one()) {
                mma_tile(a_smem, ring_addr + slot * CH_TILE, tmem_base + acc * 256 + j * 128, kb == 0);
                umma_commit(&empty[slot]);
                if (j == op.cn - 1 && a_slot >= 0) umma_commit(&empty[a_slot]);
                if (j == op.cn - 1 && kb == op.nkb - 1) umma_commit(&acc_full[acc]);
              }
              __syncwarp();
              if (++slot == CH_SLOTS) { slot = 0; ph ^= 1; }
            }
          }
        }
      } else {   // CH_FFN: A = ACT0 (x), hidden chunks through the halves of accumulator 1 / ACT1, Out in accumulator 0
        const int nc = op.ff_chunks;
        mbar_wait(&acc_empty[0], (acc_use[0] & 1) ^ 1);
        ++acc_use[0];
        mbar_wait(&acc_empty[1], (acc_use[1] & 1) ^ 1);   // accumulator 1 is only borrowed (its halves have their own barriers)
        wait_a(op);
        auto issue_h = [&](int c) {
          const uint32_t n = (hbase + c) >> 1;
          if (n > 0) mbar_wait(&hacc_empty[c & 1], (n - 1) & 1);   // the epilogue has drained the previous chunk of this half
          tc_fence_after();
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(&full[slot], ph);
            tc_fence_after();
            if (elect_one()) {
              mma_tile(act_addr + kb * CH_TILE, ring_addr + slot * CH_TILE, tmem_base + 256 + (c & 1) * 128, kb == 0);
              umma_commit(&empty[slot]);
              if (kb == 3) umma_commit(&hacc_full[c & 1]);
            }
            __syncwarp();
            if (++slot == CH_SLOTS) { slot = 0; ph ^= 1; }
          }
        };
        issue_h(0);
        if (nc > 1) issue_h(1);
        for (int c = 0; c < nc; ++c) {
          mbar_wait(&h_full[c & 1], ((hbase + c) >> 1) & 1);   // bf16 hidden chunk c is in ACT1 half (c & 1)
          tc_fence_after();
          for (int kbl = 0; kbl < 2; ++kbl)
            for (int nb = 0; nb < 2; ++nb) {
              mbar_wait(&full[slot], ph);
              tc_fence_after();
              if (elect_one()) {
                mma_tile(act_addr + (4 + (c & 1) * 2 + kbl) * CH_TILE, ring_addr + slot * CH_TILE, tmem_base + nb * 128,
                         c == 0 && kbl == 0);
                umma_commit(&empty[slot]);
                if (kbl == 1 && nb == 1) {
                  umma_commit(&h_free[c & 1]);
                  if (c == nc - 1) umma_commit(&acc_full[0]);
                }
              }
              __syncwarp();
              if (++slot == CH_SLOTS) { slot = 0; ph ^= 1; }
            }
          if (c + 2 < nc) issue_h(c + 2);
        }
        // the last two hidden chunks are drained before accumulator 1 is used again
        for (int c = (nc >= 2 ? nc - 2 : 0); c < nc; ++c) mbar_wait(&hacc_empty[c & 1], ((hbase + c) >> 1) & 1);
        hbase += nc;
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5), thread = accumulator row
    pdl_wait();
    const int quad = warp & 3;
    EpiCtx e;
    e.lrow = quad * 32 + lane;
    e.grow = row0 + e.lrow;
    e.valid = e.grow < P.M;
    e.t = (e.valid ? e.grow : 0) % P.T;
    e.act = act;
    e.headw = headw;
    e.sine_inv = P.sine_inv;
    e.tmem_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    uint32_t acc_use[2] = {0, 0};
    uint32_t hbase = 0;
    auto act_written = [&](int blk0) {   // bf16 rows of an ACT buffer are complete: visible to the tensor core, tell the MMA warp
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&act_epi[blk0 >> 2]);
    };
    auto release_acc = [&](int acc) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    };
    for (int oi = 0; oi < P.n_ops; ++oi) {
      const ChOp& op = P.ops[oi];
      if (op.kind == CH_LOAD) continue;
      if (op.kind == CH_GEMM) {
        const int chunks = op.nblk / op.cn;
        int acc = op.acc0;
        for (int ch = 0; ch < chunks; ++ch, acc ^= 1) {
          mbar_wait(&acc_full[acc], acc_use[acc] & 1);
          ++acc_use[acc];
          tc_fence_after();
          const uint32_t tacc = e.tmem_row + acc * 256;
          if (op.ln_w != nullptr) epi_ln(op, e, tacc, op.bias);
          else epi_plain(op, e, tacc, ch * op.cn * 128, op.cn * 128);
          release_acc(acc);
        }
        if (op.out_act >= 0) act_written(op.out_act);
        else if (op.gen_sine) act_written(0);
      } else {   // CH_FFN
        const int nc = op.ff_chunks;
        uint32_t raw[32];
        float v[32];
        for (int c = 0; c < nc; ++c) {
          const uint32_t n = (hbase + c) >> 1;
          mbar_wait(&hacc_full[c & 1], n & 1);
          if (n > 0) mbar_wait(&h_free[c & 1], (n - 1) & 1);   // the second GEMM has consumed the previous chunk of this ACT1 half
          tc_fence_after();
          const uint32_t th = e.tmem_row + 256 + (c & 1) * 128;
          for (int g = 0; g < 4; ++g) {
            tmem_ld32(th + g * 32, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
            add_vec32(v, op.bias + c * 128 + g * 32);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            store_act32(e, 4 + (c & 1) * 2, g * 32, v);
          }
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { mbar_arrive(&hacc_empty[c & 1]); mbar_arrive(&h_full[c & 1]); }
        }
        hbase += nc;
        mbar_wait(&acc_full[0], acc_use[0] & 1);
        ++acc_use[0];
        tc_fence_after();
        epi_ln(op, e, e.tmem_row, op.bias2);
        release_acc(0);
        if (op.out_act >= 0) act_written(op.out_act);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host
CUtensorMap make_tmap_2d(const void* ptr, int rows, int cols, int ld, int box_rows, bool f32);

void chain_set_tmap(ChainParams& p, int i, const bf16* ptr, int rows, int cols, int ld) {
  VG_CHECK(i >= 0 && i < CH_MAX_TM, "chain: tensor map index");
  p.tm[i] = make_tmap_2d(ptr, rows, cols, ld, 128, false);
  if (p.n_tm < i + 1) p.n_tm = i + 1;
}

void chain_launch(ChainParams& p, cudaStream_t stream) {
  VG_CHECK(p.n_ops >= 1 && p.n_ops <= CH_MAX_OPS && p.M >= 1 && p.T >= 1, "chain: bad program");
  for (int k = 0; k < 64; ++k) p.sine_inv[k] = 6.283185307179586f / powf(10000.f, (float)(2 * k) / 128.f);
  static bool attr_set = false;
  if (!attr_set) {
    VG_CUDA(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM));
    attr_set = true;
  }
  const int tiles = (p.M + 127) / 128;
  launch_pdl(chain_kernel, dim3(tiles), dim3(192), (size_t)CH_SMEM, stream, p);
}

// Host-side packing: W [N, K] fp32 (nn.Linear layout) → bf16 tiles [N/128][K/64][128 rows x 64] in the 128B-swizzled K-major
// UMMA operand layout (rows beyond N are zero).  Returns the number of bf16 elements written (a multiple of 8192).
size_t chain_tile_weights(const float* W, int N, int K, std::vector<bf16>& out) {
  VG_CHECK(K % 64 == 0, "chain: K must be a multiple of 64");
  const int nblk = (N + 127) / 128, nkb = K / 64;
  out.assign((size_t)nblk * nkb * 8192, __float2bfloat16(0.f));
  for (int nb = 0; nb < nblk; ++nb)
    for (int kb = 0; kb < nkb; ++kb) {
      bf16* t = out.data() + ((size_t)nb * nkb + kb) * 8192;
      for (int r = 0; r < 128; ++r) {
        const int n = nb * 128 + r;
        if (n >= N) break;
        const float* src = W + (size_t)n * K + kb * 64;
        for (int c = 0; c < 64; ++c) t[r * 64 + (((c >> 3) ^ (r & 7)) << 3) + (c & 7)] = __float2bfloat16(src[c]);
      }
    }
  return out.size();
}

}  // namespace vg
