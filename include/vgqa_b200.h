/* vgqa_b200 — C-ABI of the B200-native (sm_100a) VGQA grounding hot path.
 *
 * This is the drop-in boundary for the path BASELINE.json's north_star names: everything
 * `VSTGNet.forward` does after the ResNet101 / Video-Swin / RoBERTa feature extractors
 * (reference vgqa/core/grounding_net.py:114-202) plus `PostProcess.forward`
 * (vgqa/core/postprocessor.py:14-50).  The reference has no FFI of its own (it is pure Python); the seam
 * is its Python factory layer (vgqa/core/decoder/__init__.py:6-20, vgqa/core/__init__.py:8-54).  The Python
 * mirror of that seam lives in vgqa_b200/modules.py and binds these symbols with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions: plain pointers and sizes only; every entry point returns 0 on success and non-zero on
 * failure, in which case vgqa_last_error() (thread-local) describes the problem — the Python side raises
 * RuntimeError / AssertionError like the reference's asserts do.  No allocation happens inside
 * vgqa_forward(): all workspace is owned by the context.  One context per device and per thread.
 */
#ifndef VGQA_B200_H_
#define VGQA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vgqa_ctx vgqa_ctx;

/* Model hyper-parameters (reference: vgqa/config/defaults.py:63-72, :7, :86-87) and workspace capacity. */
typedef struct {
  int enc_layers;    /* MODEL.VSTG.ENC_LAYERS (6) */
  int dec_layers;    /* MODEL.VSTG.DEC_LAYERS (6) */
  int hidden;        /* MODEL.VSTG.HIDDEN — must be 256 */
  int heads;         /* MODEL.VSTG.HEADS  — must be 8   */
  int ffn_dim;       /* MODEL.VSTG.FFN_DIM (2048), multiple of 256 */
  int app_num;       /* DATASET.APP_NUM (20) — vocab of s_spatial_clas */
  int mot_num;       /* DATASET.MOT_NUM (34) — vocab of t_spatial_clas */
  int max_video_len; /* INPUT.MAX_VIDEO_LEN: the `te` buffers have max_video_len+1 rows */
  int max_clips;     /* capacity: clips per forward call                                   */
  int max_frames;    /* capacity: frames per clip (T), <= max_video_len + 1                */
  int max_hw;        /* capacity: feature-map tokens per frame (H*W)                       */
  int max_text;      /* capacity: text tokens (L)                                          */
  int use_cuda_graph; /* 1: capture each shape once and replay it (the caller's pointers may change) */
} vgqa_config;

/* Inputs of one forward call: `clips` independent clips (the reference is batch-1; a batch is N independent
 * B=1 forwards, grounding_net.py:108-110).  All pointers are DEVICE pointers for vgqa_forward() and HOST
 * pointers for vgqa_forward_host().  Layouts are the reference's own (NCHW fp32):
 *   vis, vid : [clips, T, 256, H, W]  = input_proj(ResNet101 feats), input_proj2(Video-Swin feats)
 *   text     : [clips, L, 256]        = text_encoder resizer output (reference shape (L,1,256) per clip)
 *   pos      : [pos_frames, 256, H, W] PositionEmbeddingSine; pos_frames = 1 (shared by every frame — the case
 *              of all-False masks) or clips*T (per frame).  NULL: the library generates it (PositionEmbeddingSine(128,
 *              normalize=True), vgqa/core/vision/position_encoding.py:50-91, from vis_mask — one shared table when
 *              vis_mask is NULL); pos_frames is then ignored
 *   vis_mask : [clips*T, H*W] uint8 (1 = padded) or NULL; text_mask: [clips, L] uint8 or NULL
 *   ori_sizes_hw : [clips, 2] fp32 (h, w) for PostProcess box scaling, or NULL (boxes_px is then not written)
 *   force_choose1/2 : optional [clips, T] fp32 0/1 masks overriding the pass-1 / pass-2 frame selection
 *              (parity tests feed the reference's decisions through these); NULL in production.       */
typedef struct {
  int clips, T, H, W, L;
  const float* vis;
  const float* vid;
  const float* text;
  const float* pos;
  int pos_frames;
  const uint8_t* vis_mask;
  const uint8_t* text_mask;
  const float* ori_sizes_hw;
  const float* force_choose1;
  const float* force_choose2;
  int iteration_rate; /* < 0 (inference): two decoder passes; >= 0: one pass (grounding_net.py:143) */
  int stop_after_encoder; /* 1: run CrossModalEncoder only (outputs encoded_feature / frames_cls) — the
                             `build_encoder(cfg)` seam (vgqa/core/decoder/__init__.py:6-8) */
  /* Optional RAW extractor outputs.  A non-NULL pointer replaces vis / vid / text (which may then be
   * NULL) and the library applies the reference's own projection on the way into the encoder's token rows:
   *   vis_raw  [clips, T, vis_raw_ch, H, W]  ResNet101 layer-4 map   → input_proj  (Conv2d 1x1, grounding_net.py:62,101)
   *   vid_raw  [clips, T, vid_raw_ch, H, W]  Video-Swin stage-3 map  → input_proj2 (Conv2d 1x1, grounding_net.py:71,105)
   *   text_raw [clips, L, text_raw_ch]       RoBERTa last_hidden_state → text_encoder.resizer (Linear + LayerNorm 1e-12,
   *                                          vgqa/core/language/bert.py:63-96)
   * Channel counts must match the weights given to vgqa_set_weight ("input_proj.weight" [256,C,1,1], "input_proj2.weight",
   * "text_encoder.resizer.fc.weight" [256,C]) and be multiples of 64; a raw input whose weights were not set is an error. */
  const float* vis_raw;
  const float* vid_raw;
  const float* text_raw;
  int vis_raw_ch, vid_raw_ch, text_raw_ch;
  /* Optional RoBERTa token ids [clips, L] (int32, <s> ... </s>, pad id 1; L <= 64): replaces text / text_raw — the library runs
   * the text tower itself (`self.body(**tokenized).last_hidden_state`, vgqa/core/language/bert.py:49,66-69: transformers
   * RobertaModel embeddings + encoder layers, weights "text_encoder.body.*") and then text_encoder.resizer.  Padded positions
   * are given by text_mask (1 = padded = `attention_mask.ne(1)`, bert.py:70), which also masks them in the cross-modal encoder. */
  const int32_t* text_ids;
  /* Layout of vis_raw / vid_raw: 0 = the reference's NCHW fp32 ([clips, T, C, H, W]); 1 = channels-last bf16
   * ([clips, T, H, W, C] bf16 behind the same pointers — what a bf16 channels_last backbone emits on B200): the map is then the
   * K-major GEMM operand itself and is read by TMA at half the bytes, with no conversion pass. */
  int raw_layout;
  /* Layout of the PROJECTED maps vis / vid: 0 = the reference's NCHW fp32 ([clips, T, 256, H, W]); 1 = channels-last bf16
   * ([clips, T, H, W, 256] bf16 behind the same pointers): each (frame, position) is then already a token row of the encoder —
   * half the bytes over PCIe for the host entry points and no transposition pass.  (The path rounds its GEMM operands to bf16
   * anyway; only the fp32 residual stream starts from the rounded values.) */
  int feat_layout;
} vgqa_inputs;

/* Outputs (fp32 unless noted); any pointer may be NULL to skip that output.
 *   pred_boxes [clips,T,4] cxcywh in (0,1); pred_sted [clips,T,2]; pred_actioness [clips,T];
 *   logits_f_m / logits_f_a [clips,T]; logits_r_a [clips,app_num]; logits_r_m [clips,mot_num];
 *   att_sequences [clips,T]; aux_boxes [dec_layers,clips,T,4]; aux_sted [dec_layers,clips,T,2];
 *   aux_actioness [dec_layers,clips,T]; choose1 / choose2 [clips,T] 0/1 (frames selected in pass 1 / 2);
 *   actioness_pass1 [clips,T] (sigmoid, pass 1); boxes_px [clips,T,4] xyxy pixels (PostProcess);
 *   sted_idx int32 [clips,2] = argmax (start,end) frame indices (PostProcess);
 *   encoded_feature [clips*T, S, 256] fp32 (frame-major; reference layout is (S,T,256)); frames_cls [clips*T,256] */
typedef struct {
  float* pred_boxes;
  float* pred_sted;
  float* pred_actioness;
  float* logits_f_m;
  float* logits_f_a;
  float* logits_r_a;
  float* logits_r_m;
  float* att_sequences;
  float* aux_boxes;
  float* aux_sted;
  float* aux_actioness;
  float* choose1;
  float* choose2;
  float* actioness_pass1;
  float* boxes_px;
  int32_t* sted_idx;
  float* encoded_feature;
  float* frames_cls;
} vgqa_outputs;

const char* vgqa_last_error(void);

/* Lifecycle. vgqa_set_weight copies one tensor of the reference state_dict (HOST fp32, reference key name,
 * e.g. "ground_encoder.encoder.spatial_layers.0.self_attn.in_proj_weight"); unknown / unused keys are accepted
 * and ignored (load_state_dict(strict=False) semantics, vgqa/inference/grounding.py:102-120).
 * vgqa_finalize_weights packs them for the kernels (bf16, fused / absorbed layouts, constant tables) and fails
 * if a key the hot path reads is missing. */
int vgqa_create(const vgqa_config* cfg, vgqa_ctx** out);
void vgqa_destroy(vgqa_ctx* ctx);
int vgqa_set_weight(vgqa_ctx* ctx, const char* name, const float* data, const int64_t* shape, int ndim);
int vgqa_finalize_weights(vgqa_ctx* ctx);

/* The hot path on device-resident inputs/outputs, enqueued on `stream` (cudaStream_t). */
int vgqa_forward(vgqa_ctx* ctx, const vgqa_inputs* in, const vgqa_outputs* out, void* stream);
/* Pipelined variant.  The forward has two phases on internal streams — phase 0: CrossModalEncoder, phase 1: classifiers,
 * decoders, heads — and the tensors between them are double-buffered in two `slot`s, so phase 1 of call k overlaps
 * phase 0 of call k+1 when consecutive calls alternate slots.  vgqa_forward_async orders the work after everything
 * already enqueued on `stream` and returns; the outputs are complete once vgqa_forward_wait has made `stream` wait
 * (host_sync = 0) or blocked the host (host_sync = 1).  Inputs must stay valid until then.
 * vgqa_forward(ctx, in, out, stream) == vgqa_forward_async(.., slot 0, stream) + vgqa_forward_wait(ctx, 0, stream, 0). */
int vgqa_forward_async(vgqa_ctx* ctx, const vgqa_inputs* in, const vgqa_outputs* out, int slot, void* stream);
int vgqa_forward_wait(vgqa_ctx* ctx, int slot, void* stream, int host_sync);
/* Same with HOST buffers: stages through pinned memory, H2D, forward, D2H, synchronises. */
int vgqa_forward_host(vgqa_ctx* ctx, const vgqa_inputs* in, const vgqa_outputs* out);
/* Pipelined variant: enqueue uploads (own copy stream), the forward and the result downloads for `slot` (0 or 1) and
 * return; the host output buffers are valid after vgqa_forward_host_wait(ctx, slot).  Alternating the two slots overlaps
 * the H2D copy of call k+1 with the compute of call k.  Host buffers must stay alive (and should be pinned) until the wait. */
int vgqa_forward_host_async(vgqa_ctx* ctx, const vgqa_inputs* in, const vgqa_outputs* out, int slot);
int vgqa_forward_host_wait(vgqa_ctx* ctx, int slot);

/* Frame sharding of ONE long clip over `world` processes / GPUs: rank r passes its contiguous frames
 * [r*T, (r+1)*T) as a clips = 1, T = local-frames call.  Encoder, classifiers, cross-attention, FFNs and heads are
 * frame-local; the library calls `fn` for the few exchanges the path needs — op 0: all-gather of `count` elements per
 * rank (recv holds world*count, rank-major), op 1: in-place all-reduce(sum) of `count` elements; dtype 0 = bf16, 1 = fp32;
 * the call must be enqueued on `stream` (NCCL from torch.distributed in vgqa_b200/parallel.py).  Exchanges per call:
 * text-token mean (1 all-reduce), frame-selection counts (2), classifier/seed sums (4), and the temporal self-attention
 * K|V rows (one all-gather per decoder layer and decoder: 12 per pass).  Outputs are the local frames' rows; gather them
 * and run vgqa_postprocess for the segment decode. */
typedef void (*vgqa_exchange_fn)(void* user, int op, const void* send, void* recv, long long count, int dtype, void* stream);
int vgqa_set_sharding(vgqa_ctx* ctx, int rank, int world, vgqa_exchange_fn fn, void* user);

/* The same sharding with the exchanges done ON THE DEVICE over NVLink peer memory (vgqa_b200/csrc/p2p_exchange.cu) instead of the
 * callback: every rank calls _export (allocates its peer-mappable exchange buffer, returns a 64-byte cudaIpcMemHandle), the
 * ranks all-gather the handles by any host channel, every rank calls _import with the world x 64 bytes (rank-major).  From then
 * on the forward of this context is a frame shard of a `world`-rank clip whose all-gathers / all-reduces are single kernels
 * (push to every peer, release/acquire flags, local reduce) — no host callback, so it is captured into a CUDA graph like the
 * unsharded forward.  All ranks must issue the same sequence of forwards.  _error returns 1 if a wait ever timed out (≈2 s). */
int vgqa_shard_p2p_export(vgqa_ctx* ctx, int rank, int world, unsigned char* handle_out /* 64 bytes */);
int vgqa_shard_p2p_import(vgqa_ctx* ctx, int rank, int world, const unsigned char* handles /* world * 64 bytes */);
int vgqa_shard_p2p_error(vgqa_ctx* ctx);

/* PostProcess.forward (vgqa/core/postprocessor.py:14-50) on device tensors: boxes [clips,T,4] cxcywh, sted [clips,T,2],
 * sizes_hw [clips,2] → boxes_px [clips,T,4] xyxy pixels (clamped at 0), sted_idx int32 [clips,2] = argmax (start,end), start < end. */
int vgqa_postprocess(const float* boxes, const float* sted, const float* sizes_hw, float* boxes_px, int32_t* sted_idx,
                     int clips, int T, void* stream);

/* The RoBERTa text tower + resizer alone (tests / serving a query once for many clips): ids [clips, L] int32 device pointer,
 * text_mask [clips, L] uint8 (1 = padded) or NULL → hidden [clips, L, Hd] fp32 (= last_hidden_state, bf16-rounded: the rows the
 * resizer reads) and text [clips, L, 256] fp32 (= the `text` input of vgqa_forward); either output may be NULL.
 * vgqa_text_tower_hidden returns Hd (0 when the tower's weights were not given). */
int vgqa_text_tower(vgqa_ctx* ctx, const int32_t* ids, const uint8_t* text_mask, int clips, int L, float* hidden, float* text,
                    void* stream);
int vgqa_text_tower_hidden(const vgqa_ctx* ctx);

/* The LAST STAGE of the Video-Swin-T extractor (first piece): `vid.layers[3]` of VSTGNet = BasicLayer of two
 * SwinTransformerBlock3D (dim 768, 24 heads, window (8,7,7), mlp ratio 4; vgqa/core/vision/video_swin_transformer.py:176-275,337-398;
 * grounding_net.py:67-71,104-105), weights "vid.layers.3.blocks.{0,1}.*" under the reference's names (optional at
 * vgqa_finalize_weights).  x: channels-last fp32 [clips, T, H, W, 768] (what the stage-3 PatchMerging leaves, `b t h w c`),
 * H = W = 7 (224 px clips), T a multiple of 8.  out_bf16: channels-last bf16 [clips, T, H, W, 768] — exactly the vid_raw /
 * raw_layout = 1 input of vgqa_forward; out_f32: the same in fp32; either may be NULL.  Device pointers. */
int vgqa_swin_stage(vgqa_ctx* ctx, const float* x, int clips, int T, int H, int W, void* out_bf16, float* out_f32, void* stream);
/* The WHOLE Video-Swin-T extractor (`self.vid` of VSTGNet = VideoSwinTransformerBackbone, video_swin_transformer.py:626-685): PatchEmbed3D
 * with patch (1,4,4) + LayerNorm, four stages (depths 2/2/6/2, dims 96..768, window (8,7,7), shifted windows with the compute_mask
 * masks) and the PatchMerging layers between them; weights "vid.patch_embed.*", "vid.layers.{0-3}.blocks.*", "vid.downsamples.{0-2}.*".
 * frames: NCHW fp32 [clips*T, 3, R, R] (`videos.tensors`), R a multiple of 32 from 224 on, T >= 8 (map sides that are not
 * multiples of the (8,7,7) window are zero-padded for the attention half of every block, as the reference does, :205-211).  out_bf16 / out_f32: the last stage's map ('3' of the reference's output dict) channels-last [clips, T, R/32, R/32, 768]
 * (out_bf16 = the vid_raw / raw_layout = 1 input of vgqa_forward); stage_out: NULL or 4 device pointers (each may be NULL) that
 * receive the four stage outputs channels-last fp32 [clips, T, R/4/2^s, R/4/2^s, 96 * 2^s].  Device pointers. */
int vgqa_swin_backbone(vgqa_ctx* ctx, const float* frames, int clips, int T, int R, void* out_bf16, float* out_f32,
                       float* const* stage_out, void* stream);
/* The ResNet101 extractor (`self.vis_encoder[0].body` of VSTGNet: torchvision resnet101 with FrozenBatchNorm2d, layer4 output —
 * vgqa/core/vision/backbone.py:13-57,104-113; grounding_net.py:49,99); weights "vis_encoder.0.body.{conv1,bn1,layer1-4.*}" under the
 * reference's names (optional at vgqa_finalize_weights; the BN buffers weight / bias / running_mean / running_var are folded into
 * the convolutions).  frames: NCHW fp32 [n_frames, 3, R, R] (`videos.tensors`), R a multiple of 32 (at most 512).  out_bf16 / out_f32:
 * the layer4 map channels-last [n_frames, R/32, R/32, 2048] (out_bf16 = the vis_raw / raw_layout = 1 input of vgqa_forward);
 * layer_out: NULL or 4 device pointers (each may be NULL) that receive the outputs of layer1..4 channels-last fp32
 * [n_frames, R/4/2^l, R/4/2^l, 256 * 2^l].  Device pointers. */
int vgqa_resnet_backbone(vgqa_ctx* ctx, const float* frames, int n_frames, int R, void* out_bf16, float* out_f32,
                         float* const* layer_out, void* stream);

/* Counters: kernels launched by the last vgqa_forward call (graph replays count the captured launches). */
int vgqa_last_launch_count(const vgqa_ctx* ctx);
/* Number of CUDA graphs this context has captured so far.  Graphs are keyed on (phase, slot, shape, which optional inputs are
 * present) — never on the caller's pointers: a serving loop that passes fresh tensors on every call captures once per shape. */
int vgqa_graph_capture_count(const vgqa_ctx* ctx);
/* Debug (VGQA_TIMELINE=1 in the environment at vgqa_create): device timestamps, in ms since the context's first call, of the
 * last forward issued on `slot`: [encoder phase start, end, decoder phase start, end]. */
int vgqa_debug_phase_times(vgqa_ctx* ctx, int slot, float* ms4);
/* Algorithmic hot-path FLOPs per clip at (T,H,W,L) by the reference's op count (SURVEY.md §8d closed form). */
double vgqa_reference_flops(int T, int H, int W, int L, int enc_layers, int dec_layers, int ffn_dim, int passes);

/* Single-kernel entry points (unit tests; also usable as building blocks).  Device pointers, bf16 data. */
int vgqa_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, void* C, int ldc, int c_f32,
                   const float* bias, int bias_period, int bias_ld, int act, const void* mul, int ldmul,
                   const void* res, int ldres, const float* ln_w, const float* ln_b, float ln_eps, void* stream);
int vgqa_mha32(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv, void* O, int ldo, int groups,
               int Sq, int Sk, const uint8_t* kmask, float scale, void* stream);
/* Encoder per-frame self-attention over the packed in-projection output QKV[F*S,768] → AO[F*S,256];
 * use_tcgen05 = 1: TMEM/TMA kernel (S <= 128), 0: warp-MMA flash kernel (any S). */
int vgqa_enc_attn(const void* QKV, void* AO, int F, int S, const uint8_t* kmask, float scale, int use_tcgen05,
                  void* stream);
int vgqa_xattn1(const void* qt, const void* mem, long long frame_stride_rows, int F, int Mk, const void* posk,
                long long posk_fstride, const void* q2, const void* kpos, int ldkpos, long long kpos_fstride,
                const uint8_t* kmask, int ldmask, float scale, void* ctx_out, float* att_out, void* stream);
/* input_proj / input_proj2 (1x1 Conv2d, grounding_net.py:62,71,101,105) fused with the token-major re-layout:
 *   X[(f*S + tok0 + p), :] = W in[f, :, p] + bias  for f < F, p < P;   in [F, C, P] fp32 (NCHW), W [256, C] bf16, C % 64 == 0.
 * X (bf16), X32 (fp32, optional), XP = bf16(x + pos row) (optional; pos [pos_frames*S, 256] bf16, pos_frames 1 or F). */
int vgqa_input_proj(const float* in, int C, const void* W, const float* bias, const void* pos, int pos_frames, void* X,
                    float* X32, void* XP, int F, int S, int tok0, int P, void* stream);
/* The same projection from a channels-last bf16 map in [F, P, C] (raw_layout = 1). */
int vgqa_input_proj_nhwc(const void* in_bf16, int C, const void* W, const float* bias, const void* pos, int pos_frames, void* X,
                         float* X32, void* XP, int F, int S, int tok0, int P, void* stream);
/* Same op with the positional score terms supplied as an additive table sbias[F, 8, ldsb] (fp32, unscaled) — the form the
 * decoders use when `pos` is frame-invariant; runs the warp-per-frame streaming kernel (xattn_stream.cu). */
int vgqa_xattn1_bias(const void* qt, const void* mem, long long frame_stride_rows, int F, int Mk, const float* sbias, int ldsb,
                     const uint8_t* kmask, int ldmask, float scale, void* ctx_out, float* att_out, void* stream);
/* norm(x + sublayer(x)) as ONE GEMM:  y = LayerNorm_256(act(A W^T + bias) + res32) * ln_w + ln_b;  C = bf16(y), C32 = y (optional),
 * C2 = bf16(y + add2[row % period]) (optional).  A [M,K] bf16 (row stride lda), W [256,K] bf16, all output / residual row strides
 * 256.  Dispatches like the forward does (gemm_ln.cu: one CTA per 128-row tile, or CTA pairs for few rows and a long K). */
int vgqa_gemm_ln(const void* A, int lda, const void* W, int ldw, int M, int K, const float* bias, int act, const float* res32,
                 const float* ln_w, const float* ln_b, float eps, void* C, float* C32, void* C2, const void* add2, int add2_period,
                 void* stream);
/* Fused post-norm FFN block of one encoder layer (modal_encoder.py:175-177):
 *   y = LayerNorm(res32 + W2 relu(W1 x + b1) + b2);  C = bf16(y), C32 = y (fp32, optional), C2 = bf16(y + add2[row % period]) (optional)
 * x [M,256] bf16, W1 [F,256] bf16, W2 [256,F] bf16 (nn.Linear layouts), F % 128 == 0, all row strides 256.  One launch on
 * CTA pairs (tcgen05 cta_group::2); the [M,F] hidden activation never reaches HBM.  `epi_parts` is reserved (pass 0). */
int vgqa_ffn_fused(const void* X, const void* W1, const float* b1, const void* W2, const float* b2, int M, int F,
                   const float* res32, const float* ln_w, const float* ln_b, float eps, void* C, float* C32, void* C2,
                   const void* add2, int add2_period, int epi_parts, void* stream);
/* Debug: in-kernel timeline of vgqa_ffn_fused (32768 int64 values; zeros unless built with -DVGQA_FFN_PROFILE). */
void vgqa_ffn_prof_read(long long* dst);

#ifdef __cplusplus
}
#endif
#endif /* VGQA_B200_H_ */
