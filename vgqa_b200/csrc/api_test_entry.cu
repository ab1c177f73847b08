// C-ABI entry points that expose single kernels for unit tests (tests/test_gemm_gpu.py etc.).
#include <string>

#include "kernels.h"

namespace vg {
static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
const char* last_error_cstr() { return g_last_error.c_str(); }
}  // namespace vg

extern "C" {

const char* vgqa_last_error(void) { return vg::last_error_cstr(); }

// C = epilogue(A · W^T); all pointers are device pointers; see include/vgqa_b200.h.
int vgqa_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, void* C, int ldc,
                   int c_f32, const float* bias, int bias_period, int bias_ld, int act, const void* mul,
                   int ldmul, const void* res, int ldres, const float* ln_w, const float* ln_b, float ln_eps,
                   void* stream) {
  try {
    vg::GemmEpi ep;
    ep.C = C; ep.ldc = ldc; ep.c_f32 = c_f32; ep.bias = bias; ep.bias_period = bias_period; ep.bias_ld = bias_ld;
    ep.act = act; ep.mul = static_cast<const vg::bf16*>(mul); ep.ldmul = ldmul;
    ep.res = static_cast<const vg::bf16*>(res); ep.ldres = ldres; ep.ln_w = ln_w; ep.ln_b = ln_b; ep.ln_eps = ln_eps;
    vg::gemm_bf16_tn(static_cast<const vg::bf16*>(A), lda, static_cast<const vg::bf16*>(W), ldw, M, N, K, ep,
                     static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) {
    vg::set_last_error(e.what());
    return 1;
  }
}

int vgqa_gemm_ln(const void* A, int lda, const void* W, int ldw, int M, int K, const float* bias, int act, const float* res32,
                 const float* ln_w, const float* ln_b, float eps, void* C, float* C32, void* C2, const void* add2, int add2_period,
                 void* stream) {
  try {
    vg::GemmEpi ep;
    ep.C = C; ep.ldc = 256; ep.bias = bias; ep.bias_ld = 256; ep.act = act; ep.res32 = res32; ep.ldres32 = 256; ep.C32 = C32; ep.ldc32 = 256;
    ep.C2 = static_cast<vg::bf16*>(C2); ep.ldc2 = 256; ep.add2 = static_cast<const vg::bf16*>(add2); ep.add2_period = add2_period;
    ep.ln_w = ln_w; ep.ln_b = ln_b; ep.ln_eps = eps;
    vg::gemm_bf16_tn(static_cast<const vg::bf16*>(A), lda, static_cast<const vg::bf16*>(W), ldw, M, 256, K, ep,
                     static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) {
    vg::set_last_error(e.what());
    return 1;
  }
}

int vgqa_ffn_fused(const void* X, const void* W1, const float* b1, const void* W2, const float* b2, int M, int F,
                    const float* res32, const float* ln_w, const float* ln_b, float eps, void* C, float* C32, void* C2,
                    const void* add2, int add2_period, int epi_parts, void* stream) {
  try {
    vg::ffn_fused(static_cast<const vg::bf16*>(X), static_cast<const vg::bf16*>(W1), b1, static_cast<const vg::bf16*>(W2), b2,
                  M, F, res32, ln_w, ln_b, eps, static_cast<vg::bf16*>(C), C32, static_cast<vg::bf16*>(C2),
                  static_cast<const vg::bf16*>(add2), add2_period, epi_parts, static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) {
    vg::set_last_error(e.what());
    return 1;
  }
}

void vgqa_ffn_prof_read(long long* dst) { vg::ffn_prof_read(dst); }

int vgqa_mha32(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo, int groups,
               int Sq, int Sk, const uint8_t* kmask, float scale, void* stream) {
  try {
    vg::mha32(static_cast<const vg::bf16*>(Q), ldq, static_cast<const vg::bf16*>(K), ldk, static_cast<const vg::bf16*>(V),
              ldv, static_cast<vg::bf16*>(O), ldo, groups, Sq, Sk, kmask, scale, static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) {
    vg::set_last_error(e.what());
    return 1;
  }
}

int vgqa_xattn1(const void* qt, const void* mem, long long frame_stride_rows, int F, int Mk, const void* posk,
                long long posk_fstride, const void* q2, const void* kpos, int ldkpos, long long kpos_fstride,
                const uint8_t* kmask, int ldmask, float scale, void* ctx_out, float* att_out, void* stream) {
  try {
    vg::xattn1(static_cast<const vg::bf16*>(qt), static_cast<const vg::bf16*>(mem), frame_stride_rows, F, Mk,
               static_cast<const vg::bf16*>(posk), posk_fstride, static_cast<const vg::bf16*>(q2),
               static_cast<const vg::bf16*>(kpos), ldkpos, kpos_fstride, kmask, ldmask, scale,
               static_cast<vg::bf16*>(ctx_out), att_out, static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) {
    vg::set_last_error(e.what());
    return 1;
  }
}

int vgqa_xattn1_bias(const void* qt, const void* mem, long long frame_stride_rows, int F, int Mk, const float* sbias, int ldsb,
                     const uint8_t* kmask, int ldmask, float scale, void* ctx_out, float* att_out, void* stream) {
  try {
    vg::xattn1(static_cast<const vg::bf16*>(qt), static_cast<const vg::bf16*>(mem), frame_stride_rows, F, Mk, nullptr, 0,
               nullptr, nullptr, 0, 0, kmask, ldmask, scale, static_cast<vg::bf16*>(ctx_out), att_out,
               static_cast<cudaStream_t>(stream), sbias, ldsb);
    return 0;
  } catch (const std::exception& e) {
    vg::set_last_error(e.what());
    return 1;
  }
}

int vgqa_input_proj(const float* in, int C, const void* W, const float* bias, const void* pos, int pos_frames, void* X,
                    float* X32, void* XP, int F, int S, int tok0, int P, void* stream) {
  try {
    vg::input_proj(in, C, static_cast<const vg::bf16*>(W), bias, static_cast<const vg::bf16*>(pos), pos_frames,
                   static_cast<vg::bf16*>(X), X32, static_cast<vg::bf16*>(XP), F, S, tok0, P, static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) {
    vg::set_last_error(e.what());
    return 1;
  }
}

int vgqa_input_proj_nhwc(const void* in, int C, const void* W, const float* bias, const void* pos, int pos_frames, void* X,
                         float* X32, void* XP, int F, int S, int tok0, int P, void* stream) {
  try {
    vg::input_proj_nhwc(static_cast<const vg::bf16*>(in), C, static_cast<const vg::bf16*>(W), bias, static_cast<const vg::bf16*>(pos),
                        pos_frames, static_cast<vg::bf16*>(X), X32, static_cast<vg::bf16*>(XP), F, S, tok0, P,
                        static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) {
    vg::set_last_error(e.what());
    return 1;
  }
}

int vgqa_enc_attn(const void* QKV, void* AO, int F, int S, const uint8_t* kmask, float scale, int use_tcgen05,
                  void* stream) {
  try {
    const vg::bf16* q = static_cast<const vg::bf16*>(QKV);
    if (use_tcgen05)
      vg::enc_attn_tc(q, static_cast<vg::bf16*>(AO), F, S, kmask, scale, static_cast<cudaStream_t>(stream));
    else
      vg::mha32(q, 768, q + 256, 768, q + 512, 768, static_cast<vg::bf16*>(AO), 256, F, S, S, kmask, scale,
                static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) {
    vg::set_last_error(e.what());
    return 1;
  }
}

}  // extern "C"
