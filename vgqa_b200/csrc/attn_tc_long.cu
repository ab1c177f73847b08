// Per-frame encoder self-attention on tcgen05 / TMEM / TMA for frames with MORE than 128 tokens (12x12 .. 14x14 feature maps:
// S = 2·H·W + L = 352 / 412; BASELINE configs "384 px" and the reference yaml at 448 px) — the multi-tile form of attn_tc.cu.
//
// Reference: F.multi_head_attention_forward inside TransformerEncoderLayer (vgqa/core/decoder/modal_encoder.py:172).
//
// Work item = (frame, head, 128-query tile); it is processed as ceil(S/128) sub-units, one per 128-key tile, by ONE softmax
// warpgroup that keeps the running (max, sum, O row) of the online softmax in registers — O is only 32 fp32 per query row, so no
// correction pass over TMEM is needed: after each sub-unit's P·V the thread reads its 32 + 1 partial values and folds them in.
// The four warpgroups run four independent item streams; the TMA producer and the MMA issuer interleave the sub-units of the
// four streams round-robin, so the pipeline of attn_tc.cu carries over unchanged: 4 S tiles in TMEM (the partial O of a sub-unit
// overwrites the first 48 columns of its consumed S tile), separate Q|K and V rings, P·V issued three sub-units behind Q·K^T,
// row sums from a ones column of the V operand, two exponentials per MUFU op on packed bf16.  Rows / keys beyond S are
// zero-filled on load and clipped on store by the 3D tensor maps (token dimension = the true S).
#include "attn_tc.cuh"

namespace vg {

// one control warpgroup (TMA, MMA, two idle warps: registers released) + the four softmax warpgroups (setmaxnreg 112: the running
// 32-wide O row, two 32-wide chunk buffers and the packed probabilities stay in registers)
static constexpr int kAtLongThreads = 128 + kAtWgs * 128;

struct LongUnit {
  int f, h, qt, j;   // frame, head, query tile, key tile
  bool valid;
};

__global__ void __launch_bounds__(kAtLongThreads, 1)
enc_attn_tc_long_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_o,
                        const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_qk = smem;
  uint8_t* s_v = s_qk + kAtQkStages * kAtQkBytes;
  uint8_t* s_p = s_v + kAtVStages * kAtVBytes;
  uint8_t* s_ones = s_p + kAtWgs * kAtPBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ones + kAtOnesBytes);
  uint64_t* qk_full = bars;
  uint64_t* qk_empty = qk_full + kAtQkStages;
  uint64_t* v_full = qk_empty + kAtQkStages;
  uint64_t* v_empty = v_full + kAtVStages;
  uint64_t* s_full = v_empty + kAtVStages;
  uint64_t* p_full = s_full + kAtWgs;
  uint64_t* o_full = p_full + kAtWgs;
  uint64_t* t_free = o_full + kAtWgs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_free + kAtWgs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  const int nt = (S + 127) >> 7;   // 128-token tiles per frame (queries and keys)
  int my_frames = 0;
  for (int f = blockIdx.x; f < p.F * p.hsplit; f += gridDim.x) ++my_frames;
  const int NH = p.heads / p.hsplit;                      // heads per pseudo-frame (8 in the encoder; 1 in the Video-Swin stage)
  const int HS = p.hsplit;                                // pseudo-frames per frame
  const int koff = p.heads * 32, voff = p.heads * 64;     // column of the K / V slices in the packed [q | k | v] row
  const int items = my_frames * NH * nt;                  // (pseudo-frame, head, query tile)
  const int slots = ((items + kAtWgs - 1) / kAtWgs) * kAtWgs * nt;   // sub-unit slots, stream-interleaved: slot u → stream u % 4
  // slot u → sub-unit (u / 4) of stream (u % 4): item = stream + 4 * (sub / nt), key tile = sub % nt
  auto decode = [&](int u) {
    LongUnit x;
    const int wg = u % kAtWgs, sub = u / kAtWgs;
    const int it = wg + kAtWgs * (sub / nt);
    x.j = sub % nt;
    x.valid = it < items;
    const int fi = it / (NH * nt), rem = it - fi * NH * nt;
    const int pf = (int)blockIdx.x + fi * (int)gridDim.x;   // pseudo-frame → (frame, first head of its block)
    const int hl = rem / nt;
    x.f = pf / HS;
    x.h = (pf - x.f * HS) * NH + hl;
    x.qt = rem - hl * nt;
    return x;
  };

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
  if (warp >= 4) {
    for (int i = threadIdx.x - 128; i < kAtOnesBytes / 16; i += kAtWgs * 128)
      reinterpret_cast<uint4*>(s_ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int n = 0;   // valid sub-units so far: ring positions follow this counter (the MMA warp walks the same sequence)
      for (int u = 0; u < slots; ++u) {
        const LongUnit x = decode(u);
        if (!x.valid) continue;
        const int sq = n % kAtQkStages, sv = n % kAtVStages;
        mbar_wait(&qk_empty[sq], ((n / kAtQkStages) & 1) ^ 1);
        mbar_expect_tx(&qk_full[sq], kAtQkBytes);
        uint8_t* dst = s_qk + sq * kAtQkBytes;
        tma_load_3d(dst, &tm_qkv, &qk_full[sq], x.h * 32, x.qt * 128, x.f);
        tma_load_3d(dst + 8192, &tm_qkv, &qk_full[sq], koff + x.h * 32, x.j * 128, x.f);
        mbar_wait(&v_empty[sv], ((n / kAtVStages) & 1) ^ 1);
        mbar_expect_tx(&v_full[sv], kAtVBytes);
        tma_load_3d(s_v + sv * kAtVBytes, &tm_qkv, &v_full[sv], voff + x.h * 32, x.j * 128, x.f);
        ++n;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 48) | (1u << 16);  // B = [V | ones] is MN-major, N = 32 + 16
    int n_qk = 0, n_pv = 0;
    for (int u = 0; u < slots + kAtLag; ++u) {
      if (u < slots && decode(u).valid) {
        const int b = u % kAtWgs, sub = u / kAtWgs, sq = n_qk % kAtQkStages;
        mbar_wait(&qk_full[sq], (n_qk / kAtQkStages) & 1);
        mbar_wait(&t_free[b], (sub & 1) ^ 1);   // the partial O of this stream's previous sub-unit has been read
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
        ++n_qk;
      }
      const int v = u - kAtLag;
      if (v >= 0 && decode(v).valid) {
        const int b = v % kAtWgs, sub = v / kAtWgs, sv = n_pv % kAtVStages;
        mbar_wait(&v_full[sv], (n_pv / kAtVStages) & 1);
        mbar_wait(&p_full[b], sub & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t v_addr = smem_u32(s_v + sv * kAtVBytes);
          const uint32_t p_addr = smem_u32(s_p + b * kAtPBytes);
          const uint64_t vd = umma_desc_sw64_mnmajor(v_addr, smem_u32(s_ones) - v_addr);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t pd = umma_desc_sw128_kmajor(p_addr + (k >> 2) * 16384) + 2 * (k & 3);
            umma_bf16(tmem_base + b * 128, pd, vd + (uint64_t)((k * 1024) >> 4), idesc_o, k ? 1u : 0u);
          }
          umma_commit(&o_full[b]);
          umma_commit(&v_empty[sv]);
        }
        __syncwarp();
        ++n_pv;
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + output warpgroups: one item stream each =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;        // query row within the tile
    const int wg_tid = (warp - 4 - wg * 4) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    uint8_t* p_buf = s_p + wg * kAtPBytes;
    uint8_t* o_buf = p_buf;                  // O staging reuses the head of this warpgroup's P tile
    const uint32_t t_s = tmem_base + lane_base + wg * 128;
    const uint32_t sw = static_cast<uint32_t>(row & 7) << 4;
    int sub = 0;                             // sub-units this stream has processed (barrier phases)
    for (int it = wg; it < items; it += kAtWgs) {
      const int fi = it / (NH * nt), rem = it - fi * NH * nt;
      const int pf = (int)blockIdx.x + fi * (int)gridDim.x, hl = rem / nt, qt = rem - hl * nt;
      const int f = pf / HS, h = (pf - f * HS) * NH + hl;
      const uint8_t* km = p.kmask ? p.kmask + (size_t)f * S : nullptr;
      // additive score row of this query (relative position bias [+ shift mask]); rows beyond S are clipped on store
      const bf16* brow = nullptr;
      const uint8_t* ridk = nullptr;   // region ids of this window's tokens (shifted windows only)
      uint8_t ridq = 0;
      if (p.sbias != nullptr) {
        const int qrow = min(qt * 128 + row, S - 1);
        brow = p.sbias + ((size_t)h * S + qrow) * S;
        const int set = p.gset != nullptr ? (int)p.gset[f] : 0;
        if (set != 0) { ridk = p.rid + (size_t)set * 512; ridq = ridk[qrow]; }   // region-id rows are padded to 512 bytes
      }
      float m_run = -INFINITY, l_run = 0.f;
      float o_run[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o_run[i] = 0.f;
      for (int j = 0; j < nt; ++j, ++sub) {
        const uint32_t ph = sub & 1;
        const int kbase = j * 128;
        const int nkeys = S - kbase < 128 ? S - kbase : 128;   // valid keys of this tile
        mbar_wait(&s_full[wg], ph);
        tc_fence_after();
        uint32_t raw[32];
        auto load_chunk = [&](int c, float (&x)[32], bool add_bias) {
          tmem_ld32(t_s + c * 32, raw);
          tmem_ld_wait();
          if (km != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int key = c * 32 + i;
              x[i] = (key >= nkeys || km[kbase + key] != 0) ? -INFINITY : __uint_as_float(raw[i]);
            }
          } else if (c * 32 + 32 > nkeys) {
            const int nvalid = nkeys - c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = i < nvalid ? __uint_as_float(raw[i]) : -INFINITY;
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(raw[i]);
          }
          if (brow != nullptr && add_bias) {   // rows of S bf16 with S % 8 == 0: 16-byte loads on full chunks, guarded scalars on the tail chunk
            const bf16* b = brow + kbase + c * 32;
            if (c * 32 + 32 <= nkeys) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(b) + i);
                const float2 a0 = unpack_bf16(q.x), a1 = unpack_bf16(q.y), a2 = unpack_bf16(q.z), a3 = unpack_bf16(q.w);
                x[8 * i] += a0.x; x[8 * i + 1] += a0.y; x[8 * i + 2] += a1.x; x[8 * i + 3] += a1.y;
                x[8 * i + 4] += a2.x; x[8 * i + 5] += a2.y; x[8 * i + 6] += a3.x; x[8 * i + 7] += a3.y;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c * 32 + i < nkeys) x[i] += __bfloat162float(b[i]);
            }
            if (ridk != nullptr) {   // SW-MSA: keys of another region of the rolled map are masked (-100 in the reference)
              // 32 region ids of this chunk as eight words (rows of 512 bytes: aligned), compared four at a time; most chunks differ nowhere
              const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(ridk + kbase + c * 32));
              const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(ridk + kbase + c * 32) + 1);
              const uint32_t q4 = (uint32_t)ridq * 0x01010101u;
              const uint32_t wds[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
              uint32_t any = 0;
#pragma unroll
              for (int wi = 0; wi < 8; ++wi) any |= __vcmpne4(wds[wi], q4);
              if (any != 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (c * 32 + i < nkeys && ((__vcmpne4(wds[i >> 2], q4) >> ((i & 3) * 8)) & 1u)) x[i] += p.mask_add;
              }
            }
            // park the biased scores in TMEM: the exponential pass reads them back instead of fetching the bias row again
#pragma unroll
            for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(x[i]);
            tmem_st32(t_s + c * 32, raw);
          }
        };
        float x[32];
        float m_new = m_run;
#pragma unroll 1
        for (int c = 0; c * 32 < nkeys; ++c) {
          load_chunk(c, x, true);
#pragma unroll
          for (int i = 0; i < 32; ++i) m_new = fmaxf(m_new, x[i]);
        }
        if (brow != nullptr) tmem_st_wait();
        const float base = m_new == -INFINITY ? 0.f : m_new;
        const float nbase = -base * p.scale_log2e;
        if (wg_tid == 0) tma_store_wait_read<0>();  // the previous item's output store has drained the head of p_buf
        named_bar_sync(1 + wg, 128);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
          if (c * 32 < nkeys) {
            load_chunk(c, x, false);
#pragma unroll
            for (int i = 0; i < 16; ++i)
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
        // partial O of this key tile (relative to `base`) and its row sum → fold into the running row
        mbar_wait(&o_full[wg], ph);
        tc_fence_after();
        tmem_ld32(t_s, raw);
        const float sum = __uint_as_float(tmem_ld1(t_s + 32));  // P · ones
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_free[wg]);   // this S/O tile may be overwritten by the stream's next sub-unit
        const float alpha = m_run == -INFINITY ? 0.f : exp2f((m_run - base) * p.scale_log2e);
#pragma unroll
        for (int i = 0; i < 32; ++i) o_run[i] = fmaf(o_run[i], alpha, __uint_as_float(raw[i]));
        l_run = fmaf(l_run, alpha, sum);
        m_run = m_new;
      }
      const float inv = 1.f / l_run;
      {
        uint8_t* rowp = o_buf + row * 64;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(rowp + ((i ^ ((row >> 1) & 3)) << 4)) =
              make_uint4(pack_bf16(o_run[8 * i] * inv, o_run[8 * i + 1] * inv), pack_bf16(o_run[8 * i + 2] * inv, o_run[8 * i + 3] * inv),
                         pack_bf16(o_run[8 * i + 4] * inv, o_run[8 * i + 5] * inv), pack_bf16(o_run[8 * i + 6] * inv, o_run[8 * i + 7] * inv));
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + wg, 128);
      if (wg_tid == 0) {
        tma_store_3d(&tm_o, o_buf, h * 32, qt * 128, f);
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
CUtensorMap make_tmap_frames(const bf16* ptr, int F, int S, int cols, int ld);   // attn_tc.cu
int device_sm_count();

static void launch_long(const bf16* QKV, bf16* AO, AttnTcParams& p, cudaStream_t stream) {
  // few frames (one long clip, a handful of Swin windows): spread the heads of a frame over CTAs as well
  p.hsplit = p.F < 2 * device_sm_count() ? p.heads : 1;
  static bool attr_set = false;
  if (!attr_set) {
    VG_CUDA(cudaFuncSetAttribute(enc_attn_tc_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    attr_set = true;
  }
  CUtensorMap tq = make_tmap_frames(QKV, p.F, p.S, 96 * p.heads, p.ldq);
  CUtensorMap to = make_tmap_frames(AO, p.F, p.S, 32 * p.heads, p.ldo);
  const int units = p.F * p.hsplit;
  const int grid = units < device_sm_count() ? units : device_sm_count();
  enc_attn_tc_long_kernel<<<grid, kAtLongThreads, kAtSmem, stream>>>(tq, to, p);
  VG_CUDA(cudaGetLastError());
}

void enc_attn_tc_long(const bf16* QKV, bf16* AO, int F, int S, const uint8_t* kmask, float scale, cudaStream_t stream) {
  VG_CHECK(S > 128 && F > 0, "enc_attn_tc_long: S must exceed 128 (attn_tc.cu handles the single-tile case)");
  AttnTcParams p;
  p.kmask = kmask; p.S = S; p.F = F; p.scale_log2e = scale * 1.4426950408889634f;
  launch_long(QKV, AO, p, stream);
}

// Window attention with an additive score term (Video-Swin W-MSA / SW-MSA, video_swin_transformer.py:143-165): `groups` windows of
// S tokens, `heads` heads of 32; QKV [groups * S, ldq] (q | k | v in the first 96 * heads columns), AO [groups * S, ldo];
// sbias [heads][S][S] bf16 = relative position bias / scale; rid / gset / mask_add: the shift mask — see AttnTcParams.
void window_attn_tc(const bf16* QKV, int ldq, bf16* AO, int ldo, int groups, int S, int heads, const bf16* sbias, const uint8_t* rid,
                    const uint8_t* gset, float scale, cudaStream_t stream) {
  VG_CHECK(S > 128 && S <= 512 && groups > 0 && heads >= 1 && sbias != nullptr && S % 8 == 0,
           "window_attn_tc: bad arguments (windows of 129..512 tokens)");
  AttnTcParams p;
  p.kmask = nullptr; p.S = S; p.F = groups; p.scale_log2e = scale * 1.4426950408889634f;
  p.heads = heads; p.ldq = ldq; p.ldo = ldo; p.sbias = sbias; p.rid = rid; p.gset = rid != nullptr ? gset : nullptr; p.mask_add = -100.0f / scale;
  launch_long(QKV, AO, p, stream);
}

}  // namespace vg
