// Device-side exchange for a frame-sharded clip: all-gather / all-reduce over NVLink peer memory, no NCCL, no
// host callback — so the sharded forward is one capturable stream of kernels.
//
// Every rank owns one peer-mapped buffer (cudaMalloc + cudaIpc) with kChannels independent channels (one per concurrent graph
// branch of the forward; exchanges of one channel are totally ordered, different channels never interact):
//   [ control block x kChannels | recv[kChannels][2 parities][world slots][slot bytes] ]
// One exchange = ONE kernel per rank (all of them run at the same point of their streams):
//   push    my `bytes` go into slot[my rank] of EVERY rank's recv buffer (16-byte stores over NVLink, parity = seq & 1);
//           the last CTA to finish pushing publishes seq in flag[my rank] of every rank (st.release.sys)
//   wait    one thread per CTA spins until every rank's flag in MY control block has reached seq (ld.acquire.sys);
//           the spin is bounded (≈2 s): on timeout an error word is set and the kernel carries on instead of hanging the GPU
//   reduce  gather: recv_out[q] = slot[q] in rank order;  sum: recv_out = slot[0] + … + slot[world-1] (fixed order → every rank
//           computes bit-identical sums)
// The sequence number lives on the device and is advanced by the last CTA to leave, so the kernel is replayable in a CUDA
// graph.  Parity double-buffering is safe because a rank can only be one exchange ahead of the slowest one: entering exchange
// n+1's wait requires everybody's flag n+1, which is published after that rank has finished reading exchange n.
#include <cstddef>
#include <cstring>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vg {

struct P2pControl {            // at the start of every rank's exchange buffer
  unsigned long long flag[16]; // flag[q] = last sequence number rank q has pushed into this rank's buffer
  unsigned long long seq;      // sequence number of the last completed exchange on this rank
  unsigned int pushed;         // CTAs of the running kernel that have finished their pushes
  unsigned int left;           // CTAs of the running kernel that have finished the exchange
  unsigned int error;          // 1 = a wait timed out
  unsigned int pad[27];
};
static_assert(sizeof(P2pControl) == 256, "control block");

static constexpr int kP2pChannels = 4;   // three graph branches of the decoder phase + the encoder phase (other stream)

struct P2pParams {
  unsigned char* peer[16];     // peer[q] = base of rank q's exchange buffer as mapped in THIS process
  int rank, world;
  long long slot_bytes;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// op 0: gather (recv gets world * bytes), op 1: fp32 sum (recv gets bytes); send may alias recv for op 1.
__global__ void __launch_bounds__(512) p2p_exchange_kernel(const P2pParams p, const unsigned char* send, unsigned char* recv,
                                                           long long bytes, int op, int ch) {
  P2pControl* ctl = reinterpret_cast<P2pControl*>(p.peer[p.rank]) + ch;
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) s_seq = *reinterpret_cast<volatile unsigned long long*>(&ctl->seq) + 1;
  __syncthreads();
  const unsigned long long seq = s_seq;
  const long long par_off = 256 * kP2pChannels + ((long long)ch * 2 + (long long)(seq & 1)) * p.world * p.slot_bytes;
  const long long n16 = bytes >> 4;
  const int tail = (int)((bytes & 15) >> 2);   // bytes is a multiple of 4 (checked on the host): up to 3 trailing words
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  // ---- push
  for (int q = 0; q < p.world; ++q) {
    uint4* dst = reinterpret_cast<uint4*>(p.peer[q] + par_off + (long long)p.rank * p.slot_bytes);
    const uint4* src = reinterpret_cast<const uint4*>(send);
    for (long long i = tid; i < n16; i += nth) dst[i] = src[i];
    if (tid < tail) reinterpret_cast<unsigned int*>(dst + n16)[tid] = reinterpret_cast<const unsigned int*>(src + n16)[tid];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(&ctl->pushed, 1u) == gridDim.x - 1) {
      ctl->pushed = 0;
      __threadfence_system();
      for (int q = 0; q < p.world; ++q)
        st_release_sys(&(reinterpret_cast<P2pControl*>(p.peer[q]) + ch)->flag[p.rank], seq);
    }
    // ---- wait for every rank's push into my buffer (bounded)
    const long long t0 = clock64();
    for (int q = 0; q < p.world; ++q) {
      while (ld_acquire_sys(&ctl->flag[q]) < seq) {
        if (clock64() - t0 > 4000000000LL) { ctl->error = 1; break; }
      }
    }
  }
  __syncthreads();
  // ---- gather / reduce out of my receive slots
  const unsigned char* base = p.peer[p.rank] + par_off;
  if (op == 0) {
    for (int q = 0; q < p.world; ++q) {
      const uint4* src = reinterpret_cast<const uint4*>(base + (long long)q * p.slot_bytes);
      uint4* dst = reinterpret_cast<uint4*>(recv + (long long)q * bytes);
      for (long long i = tid; i < n16; i += nth) dst[i] = __ldcg(src + i);
      if (tid < tail) reinterpret_cast<unsigned int*>(dst + n16)[tid] = __ldcg(reinterpret_cast<const unsigned int*>(src + n16) + tid);
    }
  } else {
    for (long long i = tid; i < n16; i += nth) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < p.world; ++q) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(base + (long long)q * p.slot_bytes) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      reinterpret_cast<float4*>(recv)[i] = acc;
    }
    if (tid < tail) {
      float acc = 0.f;
      for (int q = 0; q < p.world; ++q) acc += __ldcg(reinterpret_cast<const float*>(base + (long long)q * p.slot_bytes + n16 * 16) + tid);
      reinterpret_cast<float*>(recv + n16 * 16)[tid] = acc;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(&ctl->left, 1u) == gridDim.x - 1) {   // last CTA out: this exchange is complete
    ctl->left = 0;
    *reinterpret_cast<volatile unsigned long long*>(&ctl->seq) = seq;
  }
}

struct P2pState {
  P2pParams prm;
  unsigned char* own = nullptr;
  size_t bytes = 0;
  bool imported = false;
};

static P2pState* state(void*& opaque) {
  if (!opaque) opaque = new P2pState();
  return static_cast<P2pState*>(opaque);
}

void p2p_export(void*& opaque, int rank, int world, long long slot_bytes, unsigned char handle_out[64]) {
  VG_CHECK(world >= 2 && world <= 16 && rank >= 0 && rank < world, "p2p exchange: 2..16 ranks");
  P2pState* s = state(opaque);
  VG_CHECK(s->own == nullptr, "p2p exchange buffer already exported");
  slot_bytes = (slot_bytes + 255) / 256 * 256;
  s->bytes = 256 * kP2pChannels + (size_t)kP2pChannels * 2 * world * slot_bytes;
  VG_CUDA(cudaMalloc(&s->own, s->bytes));
  VG_CUDA(cudaMemset(s->own, 0, s->bytes));
  VG_CUDA(cudaDeviceSynchronize());
  s->prm.rank = rank; s->prm.world = world; s->prm.slot_bytes = slot_bytes;
  for (auto& q : s->prm.peer) q = nullptr;
  s->prm.peer[rank] = s->own;
  cudaIpcMemHandle_t h;
  VG_CUDA(cudaIpcGetMemHandle(&h, s->own));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t");
  memcpy(handle_out, &h, 64);
}

void p2p_import(void*& opaque, const unsigned char* handles) {
  P2pState* s = state(opaque);
  VG_CHECK(s->own != nullptr && !s->imported, "p2p exchange: export first, import once");
  for (int q = 0; q < s->prm.world; ++q) {
    if (q == s->prm.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)q * 64, 64);
    void* ptr = nullptr;
    VG_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    s->prm.peer[q] = static_cast<unsigned char*>(ptr);
  }
  s->imported = true;
}

bool p2p_ready(void* opaque) { return opaque != nullptr && static_cast<P2pState*>(opaque)->imported; }

void p2p_exchange(void* opaque, int op, const void* send, void* recv, long long bytes, int channel, cudaStream_t st) {
  P2pState* s = static_cast<P2pState*>(opaque);
  VG_CHECK(s && s->imported, "p2p exchange is not set up");
  VG_CHECK(channel >= 0 && channel < kP2pChannels, "p2p exchange: bad channel");
  VG_CHECK(bytes > 0 && bytes % 4 == 0 && bytes <= s->prm.slot_bytes, "p2p exchange: payload must be a multiple of 4 bytes and fit a slot");
  VG_CHECK((reinterpret_cast<uintptr_t>(send) & 15) == 0 && (reinterpret_cast<uintptr_t>(recv) & 15) == 0, "p2p exchange: unaligned buffer");
  VG_CHECK(op == 1 || bytes % 16 == 0, "p2p exchange: gathered payloads must be multiples of 16 bytes");
  int grid = (int)((bytes / 16 + 511) / 512);
  if (grid < 1) grid = 1;
  if (grid > 16) grid = 16;   // all CTAs must be co-resident: they wait for one another's pushes
  p2p_exchange_kernel<<<grid, 512, 0, st>>>(s->prm, static_cast<const unsigned char*>(send), static_cast<unsigned char*>(recv), bytes, op, channel);
  VG_CUDA(cudaGetLastError());
}

int p2p_error(void* opaque) {
  P2pState* s = static_cast<P2pState*>(opaque);
  if (!s || !s->own) return 0;
  unsigned int any = 0;
  for (int ch = 0; ch < kP2pChannels; ++ch) {
    unsigned int e = 0;
    VG_CUDA(cudaMemcpy(&e, s->own + ch * sizeof(P2pControl) + offsetof(P2pControl, error), 4, cudaMemcpyDeviceToHost));
    any |= e;
  }
  return (int)any;
}

void p2p_destroy(void*& opaque) {
  if (!opaque) return;
  P2pState* s = static_cast<P2pState*>(opaque);
  for (int q = 0; q < 16; ++q)
    if (s->imported && q != s->prm.rank && s->prm.peer[q]) cudaIpcCloseMemHandle(s->prm.peer[q]);
  if (s->own) cudaFree(s->own);
  delete s;
  opaque = nullptr;
}

}  // namespace vg
