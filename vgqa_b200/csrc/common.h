// Host-side shared declarations for the vgqa_b200 CUDA library (internal; the public C-ABI is
// include/vgqa_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <utility>
#include <string>

namespace vg {

typedef __nv_bfloat16 bf16;

void set_last_error(const std::string& s);

struct Error : std::runtime_error {
  explicit Error(const std::string& s) : std::runtime_error(s) {}
};

#define VG_CHECK(cond, msg)                                                                         \
  do {                                                                                              \
    if (!(cond)) throw ::vg::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
  } while (0)

#define VG_CUDA(expr)                                                                               \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw ::vg::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + #expr + " -> " + \
                        cudaGetErrorString(_e));                                                    \
  } while (0)

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

bool pdl_enabled();   // VGQA_PDL=0 disables programmatic dependent launch (gemm_tc.cu)
// Launch with the programmatic-stream-serialization attribute (PDL): the kernel must call pdl_wait() before it touches anything
// its predecessor in the stream produced (or still reads).
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  VG_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}

// Epilogue of the tcgen05 GEMM:  v = acc + bias[(row % bias_period) * bias_ld + col];  v = act(v);
// v *= mul[row, col];  v += res[row, col] + res32[row, col];  if (ln_w) v = LayerNorm_row(v) * ln_w + ln_b;  C[row, col] = v.
struct GemmEpi {
  void* C = nullptr;
  int ldc = 0;
  int c_f32 = 0;  // 0: bf16 output, 1: fp32 output
  const float* bias = nullptr;
  int bias_period = 1;
  int bias_ld = 0;
  int act = ACT_NONE;
  const bf16* mul = nullptr;
  int ldmul = 0;
  const bf16* res = nullptr;
  int ldres = 0;
  const float* res32 = nullptr;  // fp32 residual (added like `res`)
  int ldres32 = 0;
  float* C32 = nullptr;          // optional second, fp32 copy of the output (the fp32 residual stream)
  int ldc32 = 0;
  bf16* C2 = nullptr;            // LN path only: third output  bf16(LN(v) + add2[row % add2_period])  (x + pos for Q/K)
  int ldc2 = 0;
  const bf16* add2 = nullptr;    // [add2_period, 256] bf16
  int add2_period = 1;
  const float* ln_w = nullptr;
  const float* ln_b = nullptr;
  float ln_eps = 1e-5f;
};

// C[M,N] = epilogue(A[M,K] (row-major, lda) * W[N,K]^T (row-major, ldw)).  K % 64 == 0, N % 64 == 0.
void gemm_bf16_tn(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const GemmEpi& ep,
                  cudaStream_t stream);
// Weight-stationary variant (gemm_ws.cu) — used automatically by gemm_bf16_tn when the problem qualifies.
bool gemm_ws_supported(int N, int K, const GemmEpi& e);
void gemm_ws(const bf16* A, const bf16* A2, int a_switch_col, int lda, const bf16* W, int ldw, int M, int N, int K,
             const GemmEpi& e, cudaStream_t stream);
// Residual + LayerNorm epilogue variant (gemm_ln.cu), N = 256.
bool gemm_ln_supported(int N, int K, const GemmEpi& e);
void gemm_ln(const bf16* A, int lda, const bf16* W, int ldw, int M, int K, const GemmEpi& e, cudaStream_t stream);
// Fused FFN block (ffn_fused.cu): y = LN(res32 + W2 relu(W1 x + b1) + b2) on CTA pairs; hidden activation stays on chip.
bool ffn_fused_supported(int F);
void ffn_fused(const bf16* X, const bf16* W1, const float* b1, const bf16* W2, const float* b2, int M, int F,
               const float* res32, const float* ln_w, const float* ln_b, float eps, bf16* C, float* C32, bf16* C2,
               const bf16* add2, int add2_period, int epi_parts, cudaStream_t stream);
void ffn_prof_read(long long* dst);  // debug timeline of the fused FFN kernel (zeros unless built with -DVGQA_FFN_PROFILE)
void set_sm_budget(int n);  // cap (per host thread) on the grid of persistent kernels launched next; 0 = all SMs
int gemm_launch_count();  // number of tcgen05 GEMM launches issued so far by this process

}  // namespace vg
