// Shared pieces of the tcgen05 per-frame attention kernels (attn_tc.cu: S <= 128, one key tile; attn_tc_long.cu: any S, online
// softmax over 128-key tiles): pipeline constants, UMMA descriptors of the 64B-swizzled head slices, 3D TMA, packed exp2.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace vg {

static constexpr int kAtWgs = 4;                  // softmax warpgroups = units in flight
static constexpr int kAtLag = kAtWgs - 1;         // P·V is issued this many units behind Q·K^T
static constexpr int kAtQkBytes = 2 * 8192;       // Q, K: 128 rows x 64 B each
static constexpr int kAtVBytes = 8192;            // V
static constexpr int kAtQkStages = 3;
static constexpr int kAtVStages = kAtWgs + 1;     // a V tile is held from its load until the unit's P·V has completed
static constexpr int kAtPBytes = 32768;           // P: 2 atoms of 128 rows x 128 B (its first 8 KB double as the O staging tile)
static constexpr int kAtOnesBytes = 8192;         // constant bf16 1.0 tile: extra V columns that make the MMA emit row sums
static constexpr int kAtThreads = 64 + kAtWgs * 128;
static constexpr int kAtSmem = kAtQkStages * kAtQkBytes + kAtVStages * kAtVBytes + kAtWgs * kAtPBytes + kAtOnesBytes + 1024 + 256;
static_assert(kAtSmem <= 232448, "attention smem");

struct AttnTcParams {
  const uint8_t* kmask;  // [F, S] or nullptr
  int S, F;
  float scale_log2e;
  // attn_tc_long.cu only — window attention of the Video-Swin stages (swin.cu): `heads` heads of 32 in a packed [q | k | v] row
  // (row strides ldq / ldo elements), an additive score term sbias[head][q][k] (bf16, ALREADY divided by the softmax scale, so
  // that scale * (q·k + sbias) = scale * q·k + bias: the relative position bias) and the shift mask of SW-MSA: rid[set][token] (rows of 512 bytes) =
  // region id of a window token (compute_mask), gset[group] = the set a window uses (0 = unmasked); a key whose region differs from
  // the query's gets mask_add (= -100 / scale).
  int heads = 8;
  int ldq = 768, ldo = 256;
  const bf16* sbias = nullptr;   // bf16: the table is read once per key tile by every query row — half the L2 → SM bytes of fp32
  const uint8_t* rid = nullptr;
  const uint8_t* gset = nullptr;
  float mask_add = 0.f;
  // work distribution: a CTA walks "pseudo-frames" = (frame, block of heads / hsplit heads).  hsplit = 1: a frame and its heads stay on
  // one CTA (many frames); hsplit = heads: every (frame, head) is a unit of its own, so that a few large windows still fill the GPU
  int hsplit = 1;
};

// K-major operand, rows of 64 B (32 bf16), 64B swizzle: 8-row groups 512 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw64_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;  // SWIZZLE_64B
  return d;
}
// MN-major operand (N contiguous): rows = K index, 64 B (32 bf16 of N) per row, 64B swizzle; 8-row K groups 512 B apart
// (SBO); the second MN block (columns 32..47 = the constant ones tile → row sums of P) sits LBO bytes after the first.
__device__ __forceinline__ uint64_t umma_desc_sw64_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;   // next 32-wide MN block (the ones tile)
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// two 2^x per MUFU op on packed bf16 (ex2(-inf) = +0); the result is directly the bf16x2 word stored into P
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}

}  // namespace vg
