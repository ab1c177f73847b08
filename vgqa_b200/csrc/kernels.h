// Launchers of the non-GEMM kernels (attention.cu, small.cu). Internal header.
#pragma once
#include "common.h"

namespace vg {

// ---- attention.cu
void mha32(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O, int ldo, int groups,
           int Sq, int Sk, const uint8_t* kmask, float scale, cudaStream_t stream);
void xattn1(const bf16* qt, const bf16* mem, long long frame_stride_rows, int F, int Mk, const bf16* posk,
            long long posk_fstride, const bf16* q2, const bf16* kpos, int ldkpos, long long kpos_fstride,
            const uint8_t* kmask, int ldmask, float scale, bf16* ctx, float* att, cudaStream_t stream,
            const float* sbias = nullptr, int ldsb = 0);
// ---- xattn_stream.cu: warp-per-frame streaming form of xattn1 (no in-kernel positional terms; they arrive as sbias)
bool xattn_stream_supported(int Mk, long long frame_stride_rows);
void xattn_stream(const bf16* qt, const bf16* mem, long long frame_stride_rows, int F, int Mk, const float* sbias, int ldsb,
                  const uint8_t* kmask, int ldmask, float scale, bf16* ctx, float* att, cudaStream_t stream);
// frame-invariant kpos table → block-diagonal GEMM operand: out[l][h*Mpad + m][h*32 + d] = kposb[m][l*256 + h*32 + d]
void build_kpos_blockdiag(const bf16* kposb, int ldk, bf16* out, int layers, int Mk, int Mpad, cudaStream_t st);
// dst[r, :] = r < rows ? src[r, :] : 0 for r < rows_pad (256 bf16 columns)
void pad_rows_bf16(const bf16* src, bf16* dst, int rows, int rows_pad, cudaStream_t st);

// ---- attn_tc.cu: tcgen05 per-frame self-attention over the packed QKV buffer (S <= 128)
bool enc_attn_tc_supported(int S);
void enc_attn_tc(const bf16* QKV, bf16* AO, int F, int S, const uint8_t* kmask, float scale, cudaStream_t stream);

// ---- attn_tc_long.cu: window attention with an additive score term (Video-Swin stage, swin.cu)
void window_attn_tc(const bf16* QKV, int ldq, bf16* AO, int ldo, int groups, int S, int heads, const bf16* sbias, const uint8_t* rid,
                    const uint8_t* gset, float scale, cudaStream_t stream);

// ---- input_proj.cu: 1x1-conv projection of the extractor feature maps straight into the encoder's token rows
bool input_proj_supported(int C);
void input_proj(const float* in, int C, const bf16* W, const float* bias, const bf16* pos, int pos_frames, bf16* X, float* X32,
                bf16* XP, int F, int S, int tok0, int P, cudaStream_t stream);
void input_proj_nhwc(const bf16* in, int C, const bf16* W, const float* bias, const bf16* pos, int pos_frames, bf16* X, float* X32,
                     bf16* XP, int F, int S, int tok0, int P, cudaStream_t stream);
void f32_to_bf16(const float* in, bf16* out, size_t n, cudaStream_t st);

// ---- text_tower.cu: SIMT pieces of the RoBERTa text tower (the Linear layers are tcgen05 GEMMs)
void roberta_embed_ln(const int* ids, const float* word, const float* pos, const float* type0, const float* w, const float* b,
                      float eps, float* x32, bf16* x, int rows, int L, int Hd, int vocab, int max_pos, int pad_id, cudaStream_t st);
void ln_rows_wide(const float* x, const float* w, const float* b, float eps, float* y32, bf16* y, int rows, int Hd, cudaStream_t st);
void text_attn(const bf16* QKV, const uint8_t* pad, bf16* ctx, int Q, int L, int Hd, cudaStream_t st);

// ---- small.cu
// NCHW fp32 features → token-major bf16 rows: X[(f*S + tok0 + p), c] = in[f, c, p]   (in may be broadcast: fstride 0)
void nchw_to_tokens(const float* in, long long in_fstride, bf16* X, float* X32, const float* pos, long long pos_fstride,
                    bf16* XP, int F, int S, int tok0, int P, cudaStream_t st);
// channels-last bf16 features [F, P, 256] → token rows: X[(f*S + tok0 + p), :] = in[f, p, :] (+ fp32 copy, + bf16(x + pos row))
void rows_to_tokens(const bf16* in, const bf16* pos_tokens, int pos_frames, bf16* X, float* X32, bf16* XP, int F, int S, int tok0,
                    int P, cudaStream_t st);
// PositionEmbeddingSine(128, normalize=True) of `frames` (H, W) masks (uint8, 1 = padded; nullptr = nothing padded) → [frames,256,H,W]
void pos_sine(const uint8_t* mask, float* out, int frames, int H, int W, cudaStream_t st);
// X[(f*S + tok0 + l), :] = text[(f / T), l, :]   (text == nullptr → zeros)
void text_to_tokens(const float* text, bf16* X, float* X32, bf16* XP, int F, int T, int S, int tok0, int L,
                    cudaStream_t st);
// encoded_mask[f, :] = [vis_mask (with [f,0]=0) | text_mask[f/T] | vis_mask]
void build_encoded_mask(const uint8_t* vis_mask, const uint8_t* text_mask, uint8_t* out, int F, int T, int P, int L,
                        cudaStream_t st);
// Xf = LayerNorm(X); frames_cls[f] = mean_s Xf; pool_vis/vid[f] = mean over the P vis / vid tokens
void enc_finalize(const float* X32, const float* w, const float* b, float eps, bf16* Xf, float* frames_cls,
                  bf16* pool_vis, bf16* pool_vid, float* pool_vis32, float* pool_vid32, int F, int S, int P, int L,
                  cudaStream_t st);
// ftext[b,l,:] = mean_t Xf[(b*T+t)*S + P + l, :];  q0[f,:] = ftext[f/T, 0, :]
void text_sum(const bf16* Xf, float* sums, int B, int T, int S, int P, int L, cudaStream_t st);
void text_finish(const float* sums, float inv_t_global, bf16* ftext, bf16* q0, float* q0_32, int B, int T, int L,
                 cudaStream_t st);
// y[r, j] = act(x[r,:256] · w[j,:] + b[j]), j < N <= 64; act: 0 none, 1 sigmoid
void rowvec_head(const bf16* x, int ldx, const float* w, const float* b, float* y, int ldy, int rows, int N, int act,
                 cudaStream_t st);
void select_pass1(const float* lfm, const float* lfa, float theta, const float* force_w, float* att, float* w,
                  float* K, int B, int T, cudaStream_t st);
void select_pass2(const float* act_sigmoid, const float* force_w, float* w, float* K, int B, int T, cudaStream_t st);
// applies the "no frame selected → every frame" fall-back with the (possibly all-reduced) count K
void select_finish(float* w, float* K, int B, int T, int T_global, cudaStream_t st);
// part[f, c] = w[f] * sum_p att[f,p] * Xf[(f*S + tok0 + p), c]
void seed_partial(const bf16* Xf, const float* att, const float* w, float* part, int F, int S, int tok0, int P,
                  cudaStream_t st);
// red[b] = [sum_t w x[(b,t), 0..N) (64 slots) | sum_t part[(b,t), 0..256)]
void masked_sums(const float* logit_rows, int ldx, int N, const float* part, const float* w, float* red, int B, int T,
                 cudaStream_t st);
// logits_r[b] = red[b][:N] / K[b];  q[b] = red[b][64:] / (K[b] * P);  tgt[(b,t), :] = q[b]
void seed_finish(const float* red, const float* K, float* logits_r, int N, float* q, bf16* tgt, int ldt, float* tgt32,
                 int B, int T, int P, cudaStream_t st);
// boxes[f] = sigmoid(BertLN4(relu(W · BertLN256(frames_cls[f]) + b)))
void pos_fc_boxes(const float* frames_cls, const float* ln0w, const float* ln0b, const float* W, const float* b,
                  const float* ln4w, const float* ln4b, float* boxes, int F, cudaStream_t st);
void sine_embed(const float* boxes, bf16* sine, int F, cudaStream_t st);
void ln_rows(const float* x, int ldx, const float* w, const float* b, float eps, bf16* y, int ldy, int rows,
             cudaStream_t st);
void copy_cols_bf16(const bf16* src, int lds, bf16* dst, int ldd, int rows, int cols, cudaStream_t st);
// PostProcess (postprocessor.py:14-50): boxes → xyxy px (clamped), (start,end) = argmax_{s<e} ls[s] + le[e]
void postprocess(const float* boxes, const float* sted, const float* sizes_hw, float* boxes_px, int* sted_idx, int B,
                 int T, cudaStream_t st);

// ---- p2p_exchange.cu: device-side all-gather / all-reduce of a frame-sharded clip over NVLink peer memory
void p2p_export(void*& state, int rank, int world, long long slot_bytes, unsigned char handle_out[64]);
void p2p_import(void*& state, const unsigned char* handles);   // world x 64 bytes, rank-major
bool p2p_ready(void* state);
void p2p_exchange(void* state, int op, const void* send, void* recv, long long bytes, int channel, cudaStream_t st);   // op 0 gather, 1 fp32 sum; channel 0..3
int p2p_error(void* state);
void p2p_destroy(void*& state);

}  // namespace vg
