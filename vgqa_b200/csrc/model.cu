// Context, weight packing and the forward orchestration of the grounding hot path.
//
// Restates VSTGNet.forward lines 114-202 (vgqa/core/grounding_net.py) + PostProcess on top of the kernels in
// gemm_tc.cu / attention.cu / small.cu.  One launch sequence serves a whole batch of clips: every Linear is a
// tcgen05 GEMM over (clips x frames [x tokens]) rows, the reference's host syncs (`nonzero().tolist()`,
// grounding_net.py:126-128,144-150; `.item()`, postprocessor.py:44) are replaced by on-device 0/1 frame
// weights, and the sequence is captured once into a CUDA graph per shape.
//
// Algebraic restructuring done at weight-pack time (fp32/fp64 on the host, then rounded to bf16):
//   * positional adds are folded into fp32 bias tables: (x + pos) Wqk = x Wqk + (pos Wqk)     [encoder]
//     and (tgt + te) Wqk = tgt Wqk + (te Wqk)                                                [TimeDecoder]
//   * PosDecoder self-attention: the seven sa_* projections are multiplied into nn.MultiheadAttention's
//     in_proj (query_decoder.py:282-294) → one K=512 GEMM over [tgt | query_pos] plus a te-table
//   * every one-query-per-frame cross-attention has its key projection absorbed into the query
//     (q~_h = Wk_h^T q_h) and its value projection into the output projection (W_vo = Wo · blockdiag(Wv_h)),
//     so the memory-side Linear layers are never executed (attention.cu: xattn1).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <map>
#include <stdlib.h>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/vgqa_b200.h"
#include "chain.h"
#include "kernels.h"
#include "resnet.h"
#include "swin.h"

namespace vg {

const char* last_error_cstr();

// ------------------------------------------------------------------------------------------------ host tensors
struct HostT {
  std::vector<int64_t> shape;
  std::vector<float> v;
  int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

static void parallel_for(int n, const std::function<void(int)>& fn) {
  int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
  nt = std::min(nt, n);
  if (nt <= 1) { for (int i = 0; i < n; ++i) fn(i); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&, t] { for (int i = t; i < n; i += nt) fn(i); });
  for (auto& x : th) x.join();
}

// C[m,n] = A[m,k] * B[k,n]  (row-major, fp32 in, double accumulate)
static std::vector<float> matmul(const float* A, const float* B, int m, int k, int n) {
  std::vector<float> C((size_t)m * n);
  parallel_for(m, [&](int i) {
    std::vector<double> acc(n, 0.0);
    for (int p = 0; p < k; ++p) {
      const double a = A[(size_t)i * k + p];
      const float* b = B + (size_t)p * n;
      for (int j = 0; j < n; ++j) acc[j] += a * b[j];
    }
    for (int j = 0; j < n; ++j) C[(size_t)i * n + j] = (float)acc[j];
  });
  return C;
}
// y[m] = A[m,k] x[k] (+ b)
static std::vector<float> matvec(const float* A, const float* x, const float* b, int m, int k) {
  std::vector<float> y(m);
  for (int i = 0; i < m; ++i) {
    double acc = b ? b[i] : 0.0;
    for (int p = 0; p < k; ++p) acc += (double)A[(size_t)i * k + p] * x[p];
    y[i] = (float)acc;
  }
  return y;
}

// ------------------------------------------------------------------------------------------------ device arena
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, off = 0;
  bool dry = false;  // dry run: only measure
  void init(size_t bytes) {
    VG_CUDA(cudaMalloc(&base, bytes));
    VG_CUDA(cudaMemset(base, 0, bytes));
    cap = bytes; off = 0; dry = false;
  }
  void* alloc(size_t bytes) {
    off = (off + 255) & ~size_t(255);
    if (dry) { off += bytes; return nullptr; }
    VG_CHECK(off + bytes <= cap, "device arena exhausted");
    void* p = base + off;
    off += bytes;
    return p;
  }
  template <class T> T* get(size_t n) { return static_cast<T*>(alloc(n * sizeof(T))); }
  void release() { if (base) cudaFree(base); base = nullptr; }
};

struct Lin { bf16* W = nullptr; float* b = nullptr; int N = 0, K = 0; };
struct LNp { float* w = nullptr; float* b = nullptr; };
struct EncLayer { Lin qkv, out, ff1, ff2; LNp ln1, ln2; };
struct TsLayer { Lin q, o, inter, outp; LNp ln_a, ln_o; int kv_off; };
struct SaLayer { Lin qabs, vo, inter, outp; LNp ln_a, ln_o; };
struct Head { Lin t; LNp ln; float* dw = nullptr; float* db = nullptr; int vocab = 0; };
// *_t: the same weights as 16 KB tiles of the UMMA operand layout, streamed by the fused chain kernels (chain.cu)
struct TimeLayer { Lin qkv, out, qabs, vo, ff1, ff2; LNp ln1, ln3, ln4; float* tab = nullptr; bf16 *qkv_t = nullptr, *vo_t = nullptr, *ff1_t = nullptr, *ff2_t = nullptr; };
struct PosLayer { Lin sa, sa_out, q, sine, qabs, vo, ff1, ff2; LNp ln1, ln3, ln4; float* tab_sa = nullptr; bf16 *vo_t = nullptr, *ff1_t = nullptr, *ff2_t = nullptr; };
struct Mlp2 { Lin l0; float* w1 = nullptr; float* b1 = nullptr; int n1 = 0; };
struct TextLayer { Lin qkv, out, ff1, ff2; LNp ln1, ln2; };   // one RobertaLayer (q;k;v packed into one [3*Hd, Hd] Linear)

}  // namespace vg

using namespace vg;

struct vgqa_ctx {
  vgqa_config cfg;
  int device = 0;
  bool finalized = false;
  // VGQA_FFN_FUSED=0 selects the two-kernel FFN (gemm_ws + gemm_ln) for A/B measurements; both are CUDA paths
  // VGQA_ENC_SMS: SMs the encoder-phase persistent kernels may occupy (0 = all)
  int enc_sm_budget = [] { const char* e = getenv("VGQA_ENC_SMS"); return e ? atoi(e) : 0; }();
  // VGQA_ATTN_TC=0 selects the warp-MMA flash kernel for the encoder attention (A/B measurements)
  bool use_attn_tc = [] { const char* e = getenv("VGQA_ATTN_TC"); return e == nullptr || e[0] != '0'; }();
  bool use_ffn_fused = [] { const char* e = getenv("VGQA_FFN_FUSED"); return e == nullptr || e[0] != '0'; }();
  std::unordered_map<std::string, HostT> sd;
  Arena warena, ws;
  // ---- packed weights
  std::vector<EncLayer> enc;
  LNp enc_norm;
  TsLayer ts[2][2];  // [0]=t (vid tokens), [1]=s (vis tokens)
  Lin ts_kv;         // [2 cls * 2 layers * 512, 256] over f_text_cls
  Head ts_head[2];
  SaLayer sa[2][2];
  Head sa_head[2];
  std::vector<TimeLayer> tl;
  std::vector<PosLayer> pl;
  LNp time_norm;
  Lin kpos_all;  // [dec_layers*256, 256]
  Lin rph0, rph1, qs0, qs1, bb0, bb1;
  bf16 *bb0_t = nullptr, *bb1_t = nullptr, *rph0_t = nullptr, *rph1_t = nullptr;
  float *bb2w = nullptr, *bb2b = nullptr;
  // VGQA_CHAIN=1 runs the frame-local tail of every decoder layer as ONE fused row-tile GEMM chain (chain.cu) instead of one
  // launch per Linear.  Measured (profiles/r02_chain.md): 455 → 301 launches per step, parity-identical, but a chain streams
  // 3.4 MB of weights through ONE SM per 128-row tile (≈52 GB/s with 80 KB in flight) where the per-Linear launches spread the
  // same weights over the whole GPU — 2.72 vs 1.92 ms per clip at batch 1, 13.96 vs 13.75 ms per 64-clip step.  Off by default.
  bool use_chain = [] { const char* e = getenv("VGQA_CHAIN"); return e != nullptr && e[0] == '1'; }();
  // VGQA_CHAIN_HEAD=1: only the light end of the PosDecoder tail as a chain — bbox_embed (3 Linear) → sigmoid → box sine embedding
  // → ref_point_head (2 Linear) of the NEXT layer, one launch instead of seven, 640 KB of weights per row tile.  Measured: 455 → 377
  // launches, parity-identical, still slower (13.89 vs 13.61 ms per 64-clip step, 2.09 vs 1.83 ms per clip at batch 1): inside one
  // CTA every Linear of the dependent chain costs ≈10 us (row-per-thread epilogue over 256 columns + weight latency), a graph
  // node with PDL ≈5.5 us.  Off by default.
  bool use_chain_head = [] { const char* e = getenv("VGQA_CHAIN_HEAD"); return e != nullptr && e[0] == '1'; }();
  Mlp2 temp_embed, action_embed;
  float *pfc_ln0w, *pfc_ln0b, *pfc_W, *pfc_b, *pfc_ln4w, *pfc_ln4b;
  // optional front end: input_proj / input_proj2 (1x1 convs) and the text resizer; K == 0 → not loaded
  Lin ip_vis, ip_vid, ip_text;
  LNp ip_text_ln;
  bf16 *traw = nullptr, *tproj = nullptr;   // text_raw as bf16 rows, resizer output (bf16 twin of tproj32)
  float* tproj32 = nullptr;
  // optional RoBERTa text tower (text_encoder.body.*): token ids → last_hidden_state → resizer
  std::vector<TextLayer> tt;
  float *tt_word = nullptr, *tt_pos = nullptr, *tt_type = nullptr;
  LNp tt_emb_ln;
  int tt_vocab = 0, tt_maxpos = 0, tt_hd = 0;
  float *tx32 = nullptr, *ta32 = nullptr;
  bf16 *tx = nullptr, *ta = nullptr, *tqkv = nullptr, *tctx = nullptr, *th = nullptr;
  // ---- workspace (device)
  bf16 *X, *X1, *QKV, *AO, *HID, *Xf, *pos_enc, *kposb;
  // positional score terms of the decoders' cross-attentions as GEMM outputs (frame-invariant pos only)
  bf16 *pos_pad = nullptr, *kpos_bd = nullptr;
  float *t_sb = nullptr, *p_sb = nullptr;
  float *X32, *X1_32;  // fp32 residual stream of the encoder
  bf16* XP;            // bf16(x + pos): A operand of the Q/K in-projection
  uint8_t* encmask;
  // host-path input staging, two slots so that the upload of call k+1 overlaps the compute of call k
  struct HostSlot {
    float *vis, *vid, *text, *pos, *sizes, *f1, *f2;
    float *vis_raw = nullptr, *vid_raw = nullptr, *text_raw = nullptr;
    int* ids = nullptr;
    uint8_t *vmask, *tmask;
    cudaEvent_t done = nullptr;
    bool used = false;
  } hs[2];
  cudaStream_t h2d_stream = nullptr;
  float* frames_cls;
  bf16 *pool[2], *ftext, *q0, *kv_ts;
  float *pool32[2], *q0_32, *c_h32[4], *c_a32[4];
  bf16 *c_h[4], *c_q[2], *c_ctx[2], *c_a[4], *c_i[4], *c_qabs[2], *c_ctx8[2];  // per-classifier scratch (sets 0,1: TemporalSampling; 2,3: SpatialActivation)
  float *logit_f[2], *att_seq, *w1, *w2, *K1, *K2, *attmap[2], *logit_rows[2], *logits_r[2], *part[2], *seedq[2];
  float *t_tgt32, *t_x32, *t_x2_32, *p_tgt32, *p_x32, *p_x2_32;
  bf16 *t_tgt, *t_qkv, *t_ao, *t_x, *t_qabs, *t_ctx8, *t_x2, *t_hid, *t_inter, *t_hs;
  bf16 *p_h2 = nullptr;   // scratch of the query_scale MLP (runs beside the ref_point_head MLP, which owns p_h)
  bf16 *p_cat, *p_sine, *p_h, *p_s256, *p_qkv, *p_ao, *p_q, *p_q2, *p_qabs, *p_ctx8, *p_x2, *p_hid, *p_b1, *p_b2;
  float *boxes0, *anchors, *sted_all, *act_all, *act1, *boxes_px;
  int* sted_idx;
  // staging for vgqa_forward_host
  uint8_t* pinned = nullptr;
  size_t pinned_bytes = 0;
  cudaStream_t host_stream = nullptr;
  // The forward is split in two phases that run on their own streams: phase 0 = CrossModalEncoder (saturates the GPU),
  // phase 1 = classifiers + decoders + heads (hundreds of small launches).  The tensors that cross the boundary are
  // double-buffered ("slots") so that phase 1 of call k overlaps phase 0 of call k+1.
  struct Boundary {
    bf16 *Xf, *pool[2], *ftext, *q0, *pos_enc;
    float *frames_cls, *pool32[2], *q0_32;
    uint8_t* encmask;
    cudaEvent_t ev_in = nullptr, enc_done = nullptr, dec_done = nullptr;
    cudaEvent_t t_enc0 = nullptr, t_enc1 = nullptr, t_dec0 = nullptr, t_dec1 = nullptr;   // VGQA_TIMELINE=1: phase timestamps
    bool used = false;
  } bd[2];
  cudaEvent_t t_base = nullptr;
  bool timeline = [] { const char* e = getenv("VGQA_TIMELINE"); return e != nullptr && e[0] == '1'; }();
  cudaStream_t enc_stream = nullptr, dec_stream = nullptr;
  cudaStream_t aux_stream = nullptr;   // second branch of the fork/join sections (classifier pairs, the two decoders)
  cudaStream_t aux2_stream = nullptr;  // side branch inside the PosDecoder chain
  cudaStream_t aux3_stream = nullptr;  // fourth branch of the classifier section (TemporalSampling x2 ‖ SpatialActivation x2)
  cudaEvent_t fj[32] = {};
  // frame sharding of one long clip over `sh_world` ranks (vgqa_set_sharding); exchanges go through the callback
  int sh_rank = 0, sh_world = 1;
  vgqa_exchange_fn sh_fn = nullptr;
  void* sh_user = nullptr;
  void* p2p = nullptr;   // device-side exchange over peer memory (p2p_exchange.cu); replaces the callback when set up
  float *text_sums, *red[2];
  bf16 *t_qkv_all, *p_qkv_all;   // all-gathered in-projection rows of the temporal self-attention
  // inputs read by the captured part of a forward are staged in context-owned buffers (per boundary slot), so that a CUDA
  // graph depends on the SHAPE of a call only — never on the caller's pointers
  struct InStage { float *sizes, *f1, *f2; } ins[2];   // read by the decoder phase: one set per boundary slot
  float *text_in = nullptr, *pos_gen = nullptr;        // read by the encoder phase only (calls are serialised on its stream)
  int* ids_in = nullptr;
  uint8_t* tmask_in = nullptr;
  // optional Video-Swin-T extractor: the whole `vid.*` module or its last stage alone (vid.layers.3.*) — swin.cu
  SwinNet swin;
  // optional ResNet101 extractor (`vis_encoder.0.body.*`) — resnet.cu
  ResNet resnet;
  // graph cache (key = phase, slot, shape and the presence flags of the optional inputs)
  struct GraphEntry { cudaGraphExec_t exec; int launches; };
  std::map<std::vector<uint64_t>, GraphEntry> graphs;
  int graph_captures = 0;
  int last_launches = 0;
  int launches = 0;  // running count inside one forward
};

namespace vg {

// ------------------------------------------------------------------------------------------------ packing helpers
struct Packer {
  vgqa_ctx* c;
  const HostT& get(const std::string& name, std::initializer_list<int64_t> shape = {}) {
    auto it = c->sd.find(name);
    VG_CHECK(it != c->sd.end(), "missing weight '" + name + "' (set it with vgqa_set_weight before finalize)");
    if (shape.size()) {
      std::vector<int64_t> s(shape);
      VG_CHECK(it->second.shape == s, "weight '" + name + "' has an unexpected shape");
    }
    return it->second;
  }
  float* f32(const float* v, size_t n) {
    float* d = c->warena.get<float>(n);
    VG_CUDA(cudaMemcpy(d, v, n * sizeof(float), cudaMemcpyHostToDevice));
    return d;
  }
  float* f32(const std::vector<float>& v) { return f32(v.data(), v.size()); }
  bf16* b16(const float* v, size_t n) {
    std::vector<bf16> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16(v[i]);
    bf16* d = c->warena.get<bf16>(n);
    VG_CUDA(cudaMemcpy(d, h.data(), n * sizeof(bf16), cudaMemcpyHostToDevice));
    return d;
  }
  // the matrix as [N/128][K/64] tiles of 128 x 64 bf16 in the swizzled K-major operand layout (chain.cu)
  bf16* tiled(const float* W, int N, int K) {
    std::vector<bf16> h;
    chain_tile_weights(W, N, K, h);
    bf16* d = c->warena.get<bf16>(h.size());
    VG_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(bf16), cudaMemcpyHostToDevice));
    return d;
  }
  bf16* tiled(const std::string& p, int N, int K) { return tiled(get(p + ".weight", {N, K}).v.data(), N, K); }
  Lin lin(const float* W, const float* b, int N, int K) {
    Lin l; l.N = N; l.K = K; l.W = b16(W, (size_t)N * K); l.b = b ? f32(b, N) : nullptr; return l;
  }
  Lin lin(const std::string& p, int N, int K) {
    return lin(get(p + ".weight", {N, K}).v.data(), get(p + ".bias", {N}).v.data(), N, K);
  }
  LNp ln(const std::string& p, int n = 256) {
    LNp l; l.w = f32(get(p + ".weight", {n}).v); l.b = f32(get(p + ".bias", {n}).v); return l;
  }
};

// W_abs[h*256 + j, i] = sum_{c in head h} Wk[c, j] * Wq[c, i];  b_abs[h*256 + j] = sum_c Wk[c, j] * bq[c]
static void absorb_qk(const float* Wk, const float* Wq, const float* bq, int Kq, std::vector<float>& W,
                      std::vector<float>& b) {
  W.assign((size_t)2048 * Kq, 0.f);
  b.assign(2048, 0.f);
  parallel_for(2048, [&](int r) {
    const int h = r >> 8, j = r & 255;
    std::vector<double> acc(Kq, 0.0);
    double bb = 0.0;
    for (int c = h * 32; c < h * 32 + 32; ++c) {
      const double wk = Wk[(size_t)c * 256 + j];
      const float* q = Wq + (size_t)c * Kq;
      for (int i = 0; i < Kq; ++i) acc[i] += wk * q[i];
      bb += wk * bq[c];
    }
    for (int i = 0; i < Kq; ++i) W[(size_t)r * Kq + i] = (float)acc[i];
    b[r] = (float)bb;
  });
}
// W_vo[o, h*256 + j] = sum_{c in head h} Wo[o, c] * Wv[c, j];  b_vo[o] = sum_c Wo[o, c] bv[c] + bo[o]
static void absorb_vo(const float* Wo, const float* bo, const float* Wv, const float* bv, std::vector<float>& W,
                      std::vector<float>& b) {
  W.assign((size_t)256 * 2048, 0.f);
  b.assign(256, 0.f);
  parallel_for(256, [&](int o) {
    for (int h = 0; h < 8; ++h) {
      std::vector<double> acc(256, 0.0);
      for (int c = h * 32; c < h * 32 + 32; ++c) {
        const double wo = Wo[(size_t)o * 256 + c];
        const float* v = Wv + (size_t)c * 256;
        for (int j = 0; j < 256; ++j) acc[j] += wo * v[j];
      }
      for (int j = 0; j < 256; ++j) W[(size_t)o * 2048 + h * 256 + j] = (float)acc[j];
    }
    double bb = bo[o];
    for (int c = 0; c < 256; ++c) bb += (double)Wo[(size_t)o * 256 + c] * bv[c];
    b[o] = (float)bb;
  });
}

static void pack_weights(vgqa_ctx* c) {
  Packer P{c};
  const vgqa_config& cfg = c->cfg;
  const int F = cfg.ffn_dim, Tm = cfg.max_video_len + 1;
  size_t tower_bytes = 0;   // the optional text tower brings ≈330 MB of its own (fp32 embeddings + bf16 layers)
  for (const auto& kv : c->sd)
    if (kv.first.rfind("text_encoder.body.", 0) == 0) tower_bytes += (size_t)kv.second.numel() * 4 + 512;
  const bool have_swin = c->sd.count("vid.layers.3.blocks.0.attn.qkv.weight") != 0;
  const bool have_swin_full = c->sd.count("vid.patch_embed.proj.weight") != 0;
  const bool have_resnet = c->sd.count("vis_encoder.0.body.conv1.weight") != 0;
  c->warena.init(((size_t)448 << 20) + tower_bytes + (have_swin ? ((size_t)128 << 20) : 0) + (have_swin_full ? ((size_t)192 << 20) : 0) +
                 (have_resnet ? ((size_t)160 << 20) : 0));
  // ---------------- encoder (modal_encoder.py:143-178)
  c->enc.resize(cfg.enc_layers);
  for (int l = 0; l < cfg.enc_layers; ++l) {
    const std::string p = "ground_encoder.encoder.spatial_layers." + std::to_string(l) + ".";
    EncLayer& e = c->enc[l];
    const HostT& w = P.get(p + "self_attn.in_proj_weight", {768, 256});
    const HostT& b = P.get(p + "self_attn.in_proj_bias", {768});
    e.qkv = P.lin(w.v.data(), b.v.data(), 768, 256);
    e.out = P.lin(p + "self_attn.out_proj", 256, 256);
    e.ff1 = P.lin(p + "linear1", F, 256);
    e.ff2 = P.lin(p + "linear2", 256, F);
    e.ln1 = P.ln(p + "norm1");
    e.ln2 = P.ln(p + "norm2");
  }
  c->enc_norm = P.ln("ground_encoder.encoder.norm");
  // ---------------- classifiers (classifier.py, bert_module.py)
  const char* ts_names[2] = {"t_temporal_clas", "s_temporal_clas"};
  const char* sa_names[2] = {"t_spatial_clas", "s_spatial_clas"};
  auto head = [&](const std::string& p, int vocab) {
    Head h;
    h.t = P.lin(p + ".head.transform.dense", 256, 256);
    h.ln = P.ln(p + ".head.transform.LayerNorm");
    h.dw = P.f32(P.get(p + ".head.decoder.weight", {vocab, 256}).v);
    h.db = P.f32(P.get(p + ".head.bias", {vocab}).v);
    h.vocab = vocab;
    return h;
  };
  std::vector<float> kvW((size_t)2048 * 256), kvB(2048);
  for (int k = 0; k < 2; ++k) {
    for (int i = 0; i < 2; ++i) {
      const std::string p = std::string(ts_names[k]) + ".layer_ca." + std::to_string(i) + ".";
      TsLayer& t = c->ts[k][i];
      t.q = P.lin(p + "attention.self.query", 256, 256);
      t.kv_off = (k * 2 + i) * 512;
      const HostT& wk = P.get(p + "attention.self.key.weight", {256, 256});
      const HostT& wv = P.get(p + "attention.self.value.weight", {256, 256});
      std::copy(wk.v.begin(), wk.v.end(), kvW.begin() + (size_t)t.kv_off * 256);
      std::copy(wv.v.begin(), wv.v.end(), kvW.begin() + (size_t)(t.kv_off + 256) * 256);
      const HostT& bk = P.get(p + "attention.self.key.bias", {256});
      const HostT& bv = P.get(p + "attention.self.value.bias", {256});
      std::copy(bk.v.begin(), bk.v.end(), kvB.begin() + t.kv_off);
      std::copy(bv.v.begin(), bv.v.end(), kvB.begin() + t.kv_off + 256);
      t.o = P.lin(p + "attention.output.dense", 256, 256);
      t.ln_a = P.ln(p + "attention.output.LayerNorm");
      t.inter = P.lin(p + "hidden_intermediate.dense", 256, 256);
      t.outp = P.lin(p + "output.dense", 256, 256);
      t.ln_o = P.ln(p + "output.LayerNorm");
    }
    c->ts_head[k] = head(ts_names[k], 1);
  }
  c->ts_kv = P.lin(kvW.data(), kvB.data(), 2048, 256);
  std::vector<float> Wa, ba, Wv, bv;
  for (int k = 0; k < 2; ++k) {
    for (int i = 0; i < 2; ++i) {
      const std::string p = std::string(sa_names[k]) + ".layer_ca." + std::to_string(i) + ".";
      SaLayer& s = c->sa[k][i];
      absorb_qk(P.get(p + "attention.self.key.weight", {256, 256}).v.data(),
                P.get(p + "attention.self.query.weight", {256, 256}).v.data(),
                P.get(p + "attention.self.query.bias", {256}).v.data(), 256, Wa, ba);
      s.qabs = P.lin(Wa.data(), ba.data(), 2048, 256);
      absorb_vo(P.get(p + "attention.output.dense.weight", {256, 256}).v.data(),
                P.get(p + "attention.output.dense.bias", {256}).v.data(),
                P.get(p + "attention.self.value.weight", {256, 256}).v.data(),
                P.get(p + "attention.self.value.bias", {256}).v.data(), Wv, bv);
      s.vo = P.lin(Wv.data(), bv.data(), 256, 2048);
      s.ln_a = P.ln(p + "attention.output.LayerNorm");
      s.inter = P.lin(p + "hidden_intermediate.dense", 256, 256);
      s.outp = P.lin(p + "output.dense", 256, 256);
      s.ln_o = P.ln(p + "output.LayerNorm");
    }
    c->sa_head[k] = head(sa_names[k], k == 0 ? cfg.mot_num : cfg.app_num);
  }
  // ---------------- decoders (query_decoder.py)
  const std::string g = "ground_decoder.";
  const HostT& te = P.get(g + "time_embed.te", {Tm, 1, 256});
  c->tl.resize(cfg.dec_layers);
  c->pl.resize(cfg.dec_layers);
  std::vector<float> kposW((size_t)cfg.dec_layers * 256 * 256), kposB((size_t)cfg.dec_layers * 256);
  for (int l = 0; l < cfg.dec_layers; ++l) {
    {  // ---- TimeDecoderLayer (:425-486)
      const std::string p = g + "time_decoder.layers." + std::to_string(l) + ".";
      TimeLayer& t = c->tl[l];
      const HostT& w = P.get(p + "self_attn.in_proj_weight", {768, 256});
      const HostT& b = P.get(p + "self_attn.in_proj_bias", {768});
      t.qkv = P.lin(w.v.data(), nullptr, 768, 256);
      t.qkv_t = P.tiled(w.v.data(), 768, 256);
      // table[t] = [Wq;Wk] te_t + b  (v rows: bias only)   — folds `tgt + query_time` (:466)
      std::vector<float> tab((size_t)Tm * 768);
      parallel_for(Tm, [&](int ti) {
        for (int n = 0; n < 768; ++n) {
          double acc = b.v[n];
          if (n < 512)
            for (int k2 = 0; k2 < 256; ++k2) acc += (double)w.v[(size_t)n * 256 + k2] * te.v[(size_t)ti * 256 + k2];
          tab[(size_t)ti * 768 + n] = (float)acc;
        }
      });
      t.tab = P.f32(tab);
      t.out = P.lin(p + "self_attn.out_proj", 256, 256);
      t.ln1 = P.ln(p + "norm1");
      const HostT& cw = P.get(p + "cross_attn_image.in_proj_weight", {768, 256});
      const HostT& cb = P.get(p + "cross_attn_image.in_proj_bias", {768});
      absorb_qk(cw.v.data() + (size_t)256 * 256, cw.v.data(), cb.v.data(), 256, Wa, ba);
      t.qabs = P.lin(Wa.data(), ba.data(), 2048, 256);
      absorb_vo(P.get(p + "cross_attn_image.out_proj.weight", {256, 256}).v.data(),
                P.get(p + "cross_attn_image.out_proj.bias", {256}).v.data(), cw.v.data() + (size_t)512 * 256,
                cb.v.data() + 512, Wv, bv);
      t.vo = P.lin(Wv.data(), bv.data(), 256, 2048);
      t.vo_t = P.tiled(Wv.data(), 256, 2048);
      t.ln3 = P.ln(p + "norm3");
      t.ff1 = P.lin(p + "linear1", F, 256);
      t.ff2 = P.lin(p + "linear2", 256, F);
      t.ff1_t = P.tiled(p + "linear1", F, 256);
      t.ff2_t = P.tiled(p + "linear2", 256, F);
      t.ln4 = P.ln(p + "norm4");
    }
    {  // ---- PosDecoderLayer (:208-375)
      const std::string p = g + "decoder.layers." + std::to_string(l) + ".";
      PosLayer& q = c->pl[l];
      const HostT& iw = P.get(p + "self_attn.in_proj_weight", {768, 256});
      const HostT& ib = P.get(p + "self_attn.in_proj_bias", {768});
      auto W = [&](const char* n) -> const float* { return P.get(p + n + ".weight", {256, 256}).v.data(); };
      auto Bv = [&](const char* n) -> const float* { return P.get(p + n + ".bias", {256}).v.data(); };
      // Wsa [768, 512] over [tgt | query_pos]
      std::vector<float> Wsa((size_t)768 * 512, 0.f), tab((size_t)Tm * 768);
      const char* cont[3] = {"sa_qcontent_proj", "sa_kcontent_proj", "sa_v_proj"};
      const char* posn[2] = {"sa_qpos_proj", "sa_kpos_proj"};
      const char* timn[2] = {"sa_qtime_proj", "sa_ktime_proj"};
      for (int part = 0; part < 3; ++part) {
        const float* Win = iw.v.data() + (size_t)part * 256 * 256;
        std::vector<float> m1 = matmul(Win, W(cont[part]), 256, 256, 256);
        for (int r = 0; r < 256; ++r)
          std::copy(m1.begin() + (size_t)r * 256, m1.begin() + (size_t)(r + 1) * 256,
                    Wsa.begin() + (size_t)(part * 256 + r) * 512);
        std::vector<float> bsum(256);
        for (int i = 0; i < 256; ++i) bsum[i] = Bv(cont[part])[i];
        std::vector<float> mt;
        if (part < 2) {
          std::vector<float> m2 = matmul(Win, W(posn[part]), 256, 256, 256);
          for (int r = 0; r < 256; ++r)
            std::copy(m2.begin() + (size_t)r * 256, m2.begin() + (size_t)(r + 1) * 256,
                      Wsa.begin() + (size_t)(part * 256 + r) * 512 + 256);
          for (int i = 0; i < 256; ++i) bsum[i] += Bv(posn[part])[i] + Bv(timn[part])[i];
          mt = matmul(Win, W(timn[part]), 256, 256, 256);  // Win * Wtime
        }
        std::vector<float> b0 = matvec(Win, bsum.data(), ib.v.data() + part * 256, 256, 256);
        parallel_for(Tm, [&](int ti) {
          for (int n = 0; n < 256; ++n) {
            double acc = b0[n];
            if (part < 2)
              for (int k2 = 0; k2 < 256; ++k2) acc += (double)mt[(size_t)n * 256 + k2] * te.v[(size_t)ti * 256 + k2];
            tab[(size_t)ti * 768 + part * 256 + n] = (float)acc;
          }
        });
      }
      q.sa = P.lin(Wsa.data(), nullptr, 768, 512);
      q.tab_sa = P.f32(tab);
      q.sa_out = P.lin(p + "self_attn.out_proj", 256, 256);
      q.ln1 = P.ln(p + "norm1");
      // cross-attention query side
      const float* Wcq = W("ca_qcontent_proj");
      const float* bcq = Bv("ca_qcontent_proj");
      const float* Wkc = W("ca_kcontent_proj");
      if (l == 0) {
        // A = [query_pos | x] (K = 512): q = Wqpos qpos + Wcq x + (bqpos + bcq)   (:311-313)
        const float* Wqp = W("ca_qpos_proj");
        const float* bqp = Bv("ca_qpos_proj");
        std::vector<float> Wq((size_t)256 * 512), bq(256);
        for (int r = 0; r < 256; ++r) {
          std::copy(Wqp + (size_t)r * 256, Wqp + (size_t)(r + 1) * 256, Wq.begin() + (size_t)r * 512);
          std::copy(Wcq + (size_t)r * 256, Wcq + (size_t)(r + 1) * 256, Wq.begin() + (size_t)r * 512 + 256);
          bq[r] = bqp[r] + bcq[r];
        }
        q.q = P.lin(Wq.data(), bq.data(), 256, 512);
        absorb_qk(Wkc, Wq.data(), bq.data(), 512, Wa, ba);
        q.qabs = P.lin(Wa.data(), ba.data(), 2048, 512);
      } else {
        absorb_qk(Wkc, Wcq, bcq, 256, Wa, ba);
        q.qabs = P.lin(Wa.data(), ba.data(), 2048, 256);
      }
      q.sine = P.lin(p + "ca_qpos_sine_proj", 256, 256);
      absorb_vo(P.get(p + "cross_attn.out_proj.weight", {256, 256}).v.data(),
                P.get(p + "cross_attn.out_proj.bias", {256}).v.data(), W("ca_v_proj"), Bv("ca_v_proj"), Wv, bv);
      q.vo = P.lin(Wv.data(), bv.data(), 256, 2048);
      q.vo_t = P.tiled(Wv.data(), 256, 2048);
      q.ln3 = P.ln(p + "norm3");
      q.ff1 = P.lin(p + "linear1", F, 256);
      q.ff2 = P.lin(p + "linear2", 256, F);
      q.ff1_t = P.tiled(p + "linear1", F, 256);
      q.ff2_t = P.tiled(p + "linear2", 256, F);
      q.ln4 = P.ln(p + "norm4");
      std::copy(W("ca_kpos_proj"), W("ca_kpos_proj") + 65536, kposW.begin() + (size_t)l * 65536);
      std::copy(Bv("ca_kpos_proj"), Bv("ca_kpos_proj") + 256, kposB.begin() + (size_t)l * 256);
    }
  }
  c->kpos_all = P.lin(kposW.data(), kposB.data(), cfg.dec_layers * 256, 256);
  c->time_norm = P.ln(g + "time_decoder.norm");
  c->rph0 = P.lin(g + "decoder.ref_point_head.layers.0", 256, 512);
  c->rph1 = P.lin(g + "decoder.ref_point_head.layers.1", 256, 256);
  c->qs0 = P.lin(g + "decoder.query_scale.layers.0", 256, 256);
  c->qs1 = P.lin(g + "decoder.query_scale.layers.1", 256, 256);
  c->bb0 = P.lin("bbox_embed.layers.0", 256, 256);
  c->bb1 = P.lin("bbox_embed.layers.1", 256, 256);
  c->rph0_t = P.tiled(g + "decoder.ref_point_head.layers.0", 256, 512);
  c->rph1_t = P.tiled(g + "decoder.ref_point_head.layers.1", 256, 256);
  c->bb0_t = P.tiled("bbox_embed.layers.0", 256, 256);
  c->bb1_t = P.tiled("bbox_embed.layers.1", 256, 256);
  c->bb2w = P.f32(P.get("bbox_embed.layers.2.weight", {4, 256}).v);
  c->bb2b = P.f32(P.get("bbox_embed.layers.2.bias", {4}).v);
  c->temp_embed.l0 = P.lin("temp_embed.layers.0", 256, 256);
  c->temp_embed.w1 = P.f32(P.get("temp_embed.layers.1.weight", {2, 256}).v);
  c->temp_embed.b1 = P.f32(P.get("temp_embed.layers.1.bias", {2}).v);
  c->temp_embed.n1 = 2;
  c->action_embed.l0 = P.lin("action_embed.layers.0", 256, 256);
  c->action_embed.w1 = P.f32(P.get("action_embed.layers.1.weight", {1, 256}).v);
  c->action_embed.b1 = P.f32(P.get("action_embed.layers.1.bias", {1}).v);
  c->action_embed.n1 = 1;
  c->pfc_ln0w = P.f32(P.get(g + "pos_fc.0.weight", {256}).v);
  c->pfc_ln0b = P.f32(P.get(g + "pos_fc.0.bias", {256}).v);
  c->pfc_W = P.f32(P.get(g + "pos_fc.2.weight", {4, 256}).v);
  c->pfc_b = P.f32(P.get(g + "pos_fc.2.bias", {4}).v);
  c->pfc_ln4w = P.f32(P.get(g + "pos_fc.4.weight", {4}).v);
  c->pfc_ln4b = P.f32(P.get(g + "pos_fc.4.bias", {4}).v);
  // ---------------- optional front end: input_proj / input_proj2 (grounding_net.py:62,71), text resizer (bert.py:77-96)
  auto conv1x1 = [&](const std::string& name) {
    Lin l;
    auto it = c->sd.find(name + ".weight");
    if (it == c->sd.end()) return l;
    const HostT& w = it->second;
    VG_CHECK((w.shape.size() == 4 && w.shape[0] == 256 && w.shape[2] == 1 && w.shape[3] == 1) ||
                 (w.shape.size() == 2 && w.shape[0] == 256),
             "weight '" + name + ".weight' must be [256, C, 1, 1]");
    const int C = (int)w.shape[1];
    VG_CHECK(input_proj_supported(C), "'" + name + "': the input channel count must be a multiple of 64");
    return P.lin(w.v.data(), P.get(name + ".bias", {256}).v.data(), 256, C);
  };
  c->ip_vis = conv1x1("input_proj");
  c->ip_vid = conv1x1("input_proj2");
  c->ip_text = conv1x1("text_encoder.resizer.fc");
  if (c->ip_text.K > 0) c->ip_text_ln = P.ln("text_encoder.resizer.layer_norm");
  // ---------------- optional RoBERTa tower (transformers RobertaModel under `text_encoder.body`, bert.py:49)
  const std::string tb = "text_encoder.body.";
  auto wit = c->sd.find(tb + "embeddings.word_embeddings.weight");
  if (wit != c->sd.end()) {
    VG_CHECK(wit->second.shape.size() == 2, "word_embeddings.weight must be [vocab, hidden]");
    const int Hd = (int)wit->second.shape[1];
    VG_CHECK(Hd % 64 == 0 && Hd <= 1024, "text tower: the hidden size must be a multiple of 64, at most 1024");
    VG_CHECK(c->ip_text.K == Hd, "text tower: 'text_encoder.resizer.fc' must take the tower's hidden size");
    c->tt_hd = Hd; c->tt_vocab = (int)wit->second.shape[0];
    c->tt_word = P.f32(wit->second.v);
    const HostT& pe = P.get(tb + "embeddings.position_embeddings.weight");
    VG_CHECK(pe.shape.size() == 2 && pe.shape[1] == Hd, "position_embeddings.weight must be [max_pos, hidden]");
    c->tt_maxpos = (int)pe.shape[0];
    c->tt_pos = P.f32(pe.v);
    c->tt_type = P.f32(P.get(tb + "embeddings.token_type_embeddings.weight").v.data(), Hd);   // row 0 (type ids are all zero)
    c->tt_emb_ln = P.ln(tb + "embeddings.LayerNorm", Hd);
    for (int l = 0; c->sd.count(tb + "encoder.layer." + std::to_string(l) + ".attention.self.query.weight"); ++l) {
      const std::string p = tb + "encoder.layer." + std::to_string(l) + ".";
      TextLayer t;
      std::vector<float> W((size_t)3 * Hd * Hd), b((size_t)3 * Hd);
      const char* qkv[3] = {"query", "key", "value"};
      for (int k = 0; k < 3; ++k) {
        const HostT& w = P.get(p + "attention.self." + qkv[k] + ".weight", {Hd, Hd});
        const HostT& bb = P.get(p + "attention.self." + qkv[k] + ".bias", {Hd});
        std::copy(w.v.begin(), w.v.end(), W.begin() + (size_t)k * Hd * Hd);
        std::copy(bb.v.begin(), bb.v.end(), b.begin() + (size_t)k * Hd);
      }
      t.qkv = P.lin(W.data(), b.data(), 3 * Hd, Hd);
      t.out = P.lin(p + "attention.output.dense", Hd, Hd);
      t.ln1 = P.ln(p + "attention.output.LayerNorm", Hd);
      t.ff1 = P.lin(p + "intermediate.dense", 4 * Hd, Hd);
      t.ff2 = P.lin(p + "output.dense", Hd, 4 * Hd);
      t.ln2 = P.ln(p + "output.LayerNorm", Hd);
      c->tt.push_back(t);
    }
    VG_CHECK(!c->tt.empty(), "text tower: no encoder layers found under text_encoder.body.encoder.layer.*");
  }
  // ---------------- optional Video-Swin-T extractor (vid.patch_embed / layers / downsamples; video_swin_transformer.py:626-664)
  if (have_swin)
    c->swin.pack([&](const std::string& n) { return c->sd.count(n) != 0; },
                 [&](const std::string& n, std::vector<int64_t> shp) -> const float* {
                   auto it = c->sd.find(n);
                   VG_CHECK(it != c->sd.end(), "missing weight '" + n + "'");
                   VG_CHECK(it->second.shape == shp, "weight '" + n + "' has an unexpected shape");
                   return it->second.v.data();
                 },
                 [&](const float* v, size_t n) { return P.b16(v, n); }, [&](const float* v, size_t n) { return P.f32(v, n); });
  // ---------------- optional ResNet101 extractor (vis_encoder.0.body.*; backbone.py:104-113)
  if (have_resnet)
    c->resnet.pack([&](const std::string& n) { return c->sd.count(n) != 0; },
                   [&](const std::string& n, std::vector<int64_t> shp) -> const float* {
                     auto it = c->sd.find(n);
                     VG_CHECK(it != c->sd.end(), "missing weight '" + n + "'");
                     VG_CHECK(it->second.shape == shp, "weight '" + n + "' has an unexpected shape");
                     return it->second.v.data();
                   },
                   [&](const float* v, size_t n) { return P.b16(v, n); }, [&](const float* v, size_t n) { return P.f32(v, n); });
}

// ------------------------------------------------------------------------------------------------ workspace
static void carve_workspace(vgqa_ctx* c);
static void alloc_workspace(vgqa_ctx* c) {
  c->ws.dry = true; c->ws.off = 0;
  carve_workspace(c);
  const size_t bytes = c->ws.off + 4096;
  c->ws.init(bytes);
  carve_workspace(c);
}
static void carve_workspace(vgqa_ctx* c) {
  const vgqa_config& g = c->cfg;
  const size_t B = g.max_clips, T = g.max_frames, P = g.max_hw, L = g.max_text;
  const size_t F = B * T, S = 2 * P + L, R = F * S, FF = g.ffn_dim, D = g.dec_layers;
  Arena& a = c->ws;
  c->X = a.get<bf16>(R * 256); c->X1 = a.get<bf16>(R * 256); c->QKV = a.get<bf16>(R * 768); c->AO = a.get<bf16>(R * 256);
  c->HID = a.get<bf16>(R * FF);
  c->X32 = a.get<float>(R * 256); c->X1_32 = a.get<float>(R * 256); c->XP = a.get<bf16>(R * 256);
  c->kposb = a.get<bf16>(R * 1536);
  {
    const size_t Mm = P + L, Npad = (Mm + 63) / 64 * 64, Mpad = (Mm + 7) / 8 * 8;
    c->pos_pad = a.get<bf16>(Npad * 256); c->kpos_bd = a.get<bf16>(D * 8 * Mpad * 256);
    c->t_sb = a.get<float>(F * 8 * Npad); c->p_sb = a.get<float>(F * 8 * Mpad);
  }
  for (auto& b : c->bd) {
    b.Xf = a.get<bf16>(R * 256); b.pos_enc = a.get<bf16>(R * 256); b.encmask = a.get<uint8_t>(R);
    b.frames_cls = a.get<float>(F * 256); b.ftext = a.get<bf16>(B * L * 256); b.q0 = a.get<bf16>(F * 256);
    b.q0_32 = a.get<float>(F * 256);
    for (int k = 0; k < 2; ++k) { b.pool[k] = a.get<bf16>(F * 256); b.pool32[k] = a.get<float>(F * 256); }
  }
  for (auto& i : c->ins) { i.sizes = a.get<float>(B * 2); i.f1 = a.get<float>(F); i.f2 = a.get<float>(F); }
  c->text_in = a.get<float>(B * L * 256); c->ids_in = a.get<int>(B * L); c->tmask_in = a.get<uint8_t>(B * L);
  c->pos_gen = a.get<float>(F * 256 * P);   // PositionEmbeddingSine generated in the library (vgqa_inputs.pos == NULL)
  for (auto& h : c->hs) {
    h.vis = a.get<float>(F * 256 * P); h.vid = a.get<float>(F * 256 * P); h.text = a.get<float>(B * L * 256);
    h.pos = a.get<float>(F * 256 * P); h.sizes = a.get<float>(B * 2); h.f1 = a.get<float>(F); h.f2 = a.get<float>(F);
    h.vmask = a.get<uint8_t>(F * P); h.tmask = a.get<uint8_t>(B * L);
    if (c->ip_vis.K > 0) h.vis_raw = a.get<float>(F * c->ip_vis.K * P);
    if (c->ip_vid.K > 0) h.vid_raw = a.get<float>(F * c->ip_vid.K * P);
    if (c->ip_text.K > 0) h.text_raw = a.get<float>(B * L * c->ip_text.K);
  }
  if (!c->tt.empty()) {
    const size_t Hd = c->tt_hd, Rt = B * L;
    c->tx32 = a.get<float>(Rt * Hd); c->ta32 = a.get<float>(Rt * Hd); c->tx = a.get<bf16>(Rt * Hd); c->ta = a.get<bf16>(Rt * Hd);
    c->tqkv = a.get<bf16>(Rt * 3 * Hd); c->tctx = a.get<bf16>(Rt * Hd); c->th = a.get<bf16>(Rt * 4 * Hd);
    for (auto& h : c->hs) h.ids = a.get<int>(Rt);
  }
  if (c->ip_text.K > 0) {
    c->traw = a.get<bf16>(B * L * c->ip_text.K); c->tproj = a.get<bf16>(B * L * 256); c->tproj32 = a.get<float>(B * L * 256);
  }
  c->kv_ts = a.get<bf16>(B * L * 2048);
  c->text_sums = a.get<float>(B * L * 256);
  c->red[0] = a.get<float>(B * 320); c->red[1] = a.get<float>(B * 320);
  c->t_qkv_all = a.get<bf16>((size_t)(g.max_video_len + 1) * 768); c->p_qkv_all = a.get<bf16>((size_t)(g.max_video_len + 1) * 768);
  for (int k = 0; k < 4; ++k) {
    c->c_h32[k] = a.get<float>(F * 256); c->c_a32[k] = a.get<float>(F * 256);
    c->c_h[k] = a.get<bf16>(F * 256); c->c_a[k] = a.get<bf16>(F * 256); c->c_i[k] = a.get<bf16>(F * 256);
  }
  for (int k = 0; k < 2; ++k) {
    c->c_q[k] = a.get<bf16>(F * 256); c->c_ctx[k] = a.get<bf16>(F * 256);
    c->c_qabs[k] = a.get<bf16>(F * 2048); c->c_ctx8[k] = a.get<bf16>(F * 2048);
    c->logit_f[k] = a.get<float>(F); c->attmap[k] = a.get<float>(F * P); c->logit_rows[k] = a.get<float>(F * 64);
    c->logits_r[k] = a.get<float>(B * 64); c->part[k] = a.get<float>(F * 256); c->seedq[k] = a.get<float>(B * 256);
  }
  c->t_tgt32 = a.get<float>(F * 256); c->t_x32 = a.get<float>(F * 256); c->t_x2_32 = a.get<float>(F * 256);
  c->p_tgt32 = a.get<float>(F * 256); c->p_x32 = a.get<float>(F * 256); c->p_x2_32 = a.get<float>(F * 256);
  c->att_seq = a.get<float>(F); c->w1 = a.get<float>(F); c->w2 = a.get<float>(F); c->K1 = a.get<float>(B); c->K2 = a.get<float>(B);
  c->t_tgt = a.get<bf16>(F * 256); c->t_qkv = a.get<bf16>(F * 768); c->t_ao = a.get<bf16>(F * 256); c->t_x = a.get<bf16>(F * 256);
  c->t_qabs = a.get<bf16>(F * 2048); c->t_ctx8 = a.get<bf16>(F * 2048); c->t_x2 = a.get<bf16>(F * 256);
  c->t_hid = a.get<bf16>(F * FF); c->t_inter = a.get<bf16>(D * F * 256); c->t_hs = a.get<bf16>(D * F * 256);
  c->p_cat = a.get<bf16>(F * 768); c->p_sine = a.get<bf16>(F * 512); c->p_h = a.get<bf16>(F * 256); c->p_s256 = a.get<bf16>(F * 256);
  c->p_h2 = a.get<bf16>(F * 256);
  c->p_qkv = a.get<bf16>(F * 768); c->p_ao = a.get<bf16>(F * 256); c->p_q = a.get<bf16>(F * 256); c->p_q2 = a.get<bf16>(F * 256);
  c->p_qabs = a.get<bf16>(F * 2048); c->p_ctx8 = a.get<bf16>(F * 2048); c->p_x2 = a.get<bf16>(F * 256);
  c->p_hid = a.get<bf16>(F * FF); c->p_b1 = a.get<bf16>(F * 256); c->p_b2 = a.get<bf16>(F * 256);
  c->boxes0 = a.get<float>(F * 4); c->anchors = a.get<float>(D * F * 4); c->sted_all = a.get<float>(D * F * 2);
  c->act_all = a.get<float>(D * F); c->act1 = a.get<float>(F); c->boxes_px = a.get<float>(F * 4);
  c->sted_idx = a.get<int>(B * 2);
}

// point the "current" boundary tensors at one of the two slots (host-side; launches are enqueued in host order)
static void select_slot(vgqa_ctx* c, int slot) {
  vgqa_ctx::Boundary& b = c->bd[slot];
  c->Xf = b.Xf; c->pos_enc = b.pos_enc; c->encmask = b.encmask; c->frames_cls = b.frames_cls; c->ftext = b.ftext;
  c->q0 = b.q0; c->q0_32 = b.q0_32;
  for (int k = 0; k < 2; ++k) { c->pool[k] = b.pool[k]; c->pool32[k] = b.pool32[k]; }
}

// ------------------------------------------------------------------------------------------------ forward
struct Fwd {
  vgqa_ctx* c;
  cudaStream_t st;          // stream the helpers below launch on (main or aux)
  cudaStream_t main, aux;
  cudaStream_t aux2 = nullptr;   // side branch of the PosDecoder chain (query_scale MLP, absorbed-query GEMM)
  cudaStream_t aux3 = nullptr;   // with aux2: third / fourth branch of the classifier section
  int ev_i = 0;
  int phase = 1;
  int B, T, P, L, S, F, R;
  // independent sub-graphs (the two classifiers of a pair, TimeDecoder vs PosDecoder) run on two streams; under
  // stream capture this becomes two parallel branches of the CUDA graph.
  void fork() {
    if (aux == main) return;   // sharded mode: one stream, collectives stay in program order
    VG_CUDA(cudaEventRecord(c->fj[ev_i], main));
    VG_CUDA(cudaStreamWaitEvent(aux, c->fj[ev_i], 0));
    ev_i = (ev_i + 1) & 31;
  }
  void join() {
    if (aux == main) { st = main; return; }
    VG_CUDA(cudaEventRecord(c->fj[ev_i], aux));
    VG_CUDA(cudaStreamWaitEvent(main, c->fj[ev_i], 0));
    ev_i = (ev_i + 1) & 31;
    st = main;
  }
  // side branch off the current stream `from`: work enqueued on aux2 after side_fork(from) runs beside `from` until
  // side_join(from) (no-ops when everything runs on one stream)
  void side_fork(cudaStream_t from) {
    if (aux2 == from) return;
    VG_CUDA(cudaEventRecord(c->fj[ev_i], from));
    VG_CUDA(cudaStreamWaitEvent(aux2, c->fj[ev_i], 0));
    ev_i = (ev_i + 1) & 31;
  }
  void side_join(cudaStream_t into) {
    if (aux2 == into) return;
    VG_CUDA(cudaEventRecord(c->fj[ev_i], aux2));
    VG_CUDA(cudaStreamWaitEvent(into, c->fj[ev_i], 0));
    ev_i = (ev_i + 1) & 31;
  }
  // generic branch: `br` starts after everything enqueued on `from` so far / `into` continues after everything on `br`
  void branch_fork(cudaStream_t from, cudaStream_t br) {
    if (br == from) return;
    VG_CUDA(cudaEventRecord(c->fj[ev_i], from));
    VG_CUDA(cudaStreamWaitEvent(br, c->fj[ev_i], 0));
    ev_i = (ev_i + 1) & 31;
  }
  void branch_join(cudaStream_t br, cudaStream_t into) {
    if (br == into) return;
    VG_CUDA(cudaEventRecord(c->fj[ev_i], br));
    VG_CUDA(cudaStreamWaitEvent(into, c->fj[ev_i], 0));
    ev_i = (ev_i + 1) & 31;
  }
  void gemm(const bf16* A, int lda, const Lin& w, int M, const GemmEpi& ep) {
    gemm_bf16_tn(A, lda, w.W, w.K, M, w.N, w.K, ep, st);
    ++c->launches;
  }
  // C = act(A W^T + b)
  void linear(const bf16* A, int lda, const Lin& w, int M, bf16* C, int ldc, int act = ACT_NONE) {
    GemmEpi ep; ep.C = C; ep.ldc = ldc; ep.bias = w.b; ep.bias_period = 1; ep.bias_ld = w.N; ep.act = act;
    gemm(A, lda, w, M, ep);
  }
  // C (bf16) and C32 (fp32 residual stream, ld 256) = LN(res32 + act(A W^T + b))
  void linear_res_ln(const bf16* A, int lda, const Lin& w, int M, const float* res32, const LNp& ln, float eps,
                     bf16* C, int ldc, float* C32, int act = ACT_NONE) {
    GemmEpi ep; ep.C = C; ep.ldc = ldc; ep.bias = w.b; ep.bias_period = 1; ep.bias_ld = w.N; ep.act = act;
    ep.res32 = res32; ep.ldres32 = 256; ep.C32 = C32; ep.ldc32 = 256; ep.ln_w = ln.w; ep.ln_b = ln.b; ep.ln_eps = eps;
    gemm(A, lda, w, M, ep);
  }
  void count(int n = 1) { c->launches += n; }
  bool sharded() const { return c->sh_world > 1; }
  // exchange channel of the peer-memory exchange: one per graph branch, so that every channel sees one global order
  // (the encoder phase of the NEXT call runs on another stream beside the decoder phase of this one: it owns channel 3)
  int channel() const { return phase == 0 ? 3 : (st == main ? 0 : (st == aux ? 1 : 2)); }
  int T_global() const { return T * c->sh_world; }
  // in-place sum over ranks of `n` fp32 values (no-op for a single rank)
  void all_reduce_f32(float* buf, size_t n) {
    if (!sharded()) return;
    if (p2p_ready(c->p2p)) { p2p_exchange(c->p2p, 1, buf, buf, (long long)n * 4, channel(), st); ++c->launches; return; }
    c->sh_fn(c->sh_user, 1, buf, buf, (long long)n, 1, st);
  }
  // recv[world][rows_local * cols] <- every rank's send[rows_local * cols] (bf16)
  void all_gather_bf16(const bf16* send, bf16* recv, size_t elems_per_rank) {
    if (p2p_ready(c->p2p)) { p2p_exchange(c->p2p, 0, send, recv, (long long)elems_per_rank * 2, channel(), st); ++c->launches; return; }
    c->sh_fn(c->sh_user, 0, send, recv, (long long)elems_per_rank, 0, st);
  }
};

// RoBERTa encoder (transformers RobertaModel: embeddings + N x RobertaLayer, post-LayerNorm, eps 1e-5 as in roberta-base's
// config) over clips x L token ids → bf16 last_hidden_state rows in c->traw (the resizer's A operand).  bert.py:66-69.
static void run_text_tower(Fwd& f, const int* ids, const uint8_t* pad) {
  vgqa_ctx* c = f.c;
  cudaStream_t st = f.st;
  const int Hd = c->tt_hd, R = f.B * f.L;
  const float eps = 1e-5f;
  roberta_embed_ln(ids, c->tt_word, c->tt_pos, c->tt_type, c->tt_emb_ln.w, c->tt_emb_ln.b, eps, c->tx32, c->tx, R, f.L, Hd,
                   c->tt_vocab, c->tt_maxpos, /*pad_id=*/1, st);
  f.count();
  for (size_t l = 0; l < c->tt.size(); ++l) {
    TextLayer& t = c->tt[l];
    f.linear(c->tx, Hd, t.qkv, R, c->tqkv, 3 * Hd);
    text_attn(c->tqkv, pad, c->tctx, f.B, f.L, Hd, st);
    {  // attention.output: LayerNorm(dense(ctx) + x)
      GemmEpi ep; ep.C = c->ta32; ep.ldc = Hd; ep.c_f32 = 1; ep.bias = t.out.b; ep.bias_ld = Hd; ep.res32 = c->tx32; ep.ldres32 = Hd;
      f.gemm(c->tctx, Hd, t.out, R, ep);
    }
    ln_rows_wide(c->ta32, t.ln1.w, t.ln1.b, eps, c->ta32, c->ta, R, Hd, st);
    f.linear(c->ta, Hd, t.ff1, R, c->th, 4 * Hd, ACT_GELU);
    {  // output: LayerNorm(dense(gelu(...)) + attention_output)
      GemmEpi ep; ep.C = c->tx32; ep.ldc = Hd; ep.c_f32 = 1; ep.bias = t.ff2.b; ep.bias_ld = Hd; ep.res32 = c->ta32; ep.ldres32 = Hd;
      f.gemm(c->th, 4 * Hd, t.ff2, R, ep);
    }
    ln_rows_wide(c->tx32, t.ln2.w, t.ln2.b, eps, c->tx32, l + 1 == c->tt.size() ? c->traw : c->tx, R, Hd, st);
    f.count(3);
  }
}

// How the text of a call reaches the encoder: 0 = projected tokens (text_in), 1 = hidden states (bf16 rows in traw, resizer in
// the captured part), 2 = token ids (ids_in: text tower + resizer in the captured part)
static int text_kind(const vgqa_inputs& in) { return in.text_ids ? 2 : (in.text_raw ? 1 : 0); }

// Everything that READS THE CALLER'S POINTERS of the encoder phase: the positional table (given, or PositionEmbeddingSine
// generated here — vision/position_encoding.py:50-91), the token-major re-layout of the visual maps (or input_proj /
// input_proj2 on the raw extractor maps), the text staging and the key-padding mask.  Always launched eagerly; what follows
// (run_encoder) only touches context-owned memory and is what a CUDA graph captures.
static void ingest_encoder_inputs(Fwd& f, const vgqa_inputs& in, bool have_mask) {
  vgqa_ctx* c = f.c;
  cudaStream_t st = f.st;
  const int S = f.S, P = f.P, L = f.L, F = f.F;
  const int pf = in.pos_frames;
  const float* pos = in.pos;
  if (pos == nullptr) { pos_sine(in.vis_mask, c->pos_gen, pf, in.H, in.W, st); pos = c->pos_gen; f.count(); }
  // tokens: [vis | text | vid] per frame (modal_encoder.py:64)
  const long long pos_fs = pf > 1 ? (long long)256 * P : 0;
  // positional rows: [pos | 0 | pos] (modal_encoder.py:66)
  nchw_to_tokens(pos, (long long)256 * P, c->pos_enc, nullptr, nullptr, 0, nullptr, pf, S, 0, P, st);
  text_to_tokens(nullptr, c->pos_enc, nullptr, nullptr, pf, 1, S, P, L, st);
  nchw_to_tokens(pos, (long long)256 * P, c->pos_enc, nullptr, nullptr, 0, nullptr, pf, S, P + L, P, st);
  f.count(3);
  // visual tokens: already-projected maps, or the raw extractor maps through input_proj / input_proj2 (input_proj.cu)
  if (in.vis_raw != nullptr && in.raw_layout == 1)
    input_proj_nhwc(reinterpret_cast<const bf16*>(in.vis_raw), c->ip_vis.K, c->ip_vis.W, c->ip_vis.b, c->pos_enc, pf, c->X, c->X32,
                    c->XP, F, S, 0, P, st);
  else if (in.vis_raw != nullptr)
    input_proj(in.vis_raw, c->ip_vis.K, c->ip_vis.W, c->ip_vis.b, c->pos_enc, pf, c->X, c->X32, c->XP, F, S, 0, P, st);
  else if (in.feat_layout == 1)
    rows_to_tokens(reinterpret_cast<const bf16*>(in.vis), c->pos_enc, pf, c->X, c->X32, c->XP, F, S, 0, P, st);
  else
    nchw_to_tokens(in.vis, (long long)256 * P, c->X, c->X32, pos, pos_fs, c->XP, F, S, 0, P, st);
  if (in.vid_raw != nullptr && in.raw_layout == 1)
    input_proj_nhwc(reinterpret_cast<const bf16*>(in.vid_raw), c->ip_vid.K, c->ip_vid.W, c->ip_vid.b, c->pos_enc, pf, c->X, c->X32,
                    c->XP, F, S, P + L, P, st);
  else if (in.vid_raw != nullptr)
    input_proj(in.vid_raw, c->ip_vid.K, c->ip_vid.W, c->ip_vid.b, c->pos_enc, pf, c->X, c->X32, c->XP, F, S, P + L, P, st);
  else if (in.feat_layout == 1)
    rows_to_tokens(reinterpret_cast<const bf16*>(in.vid), c->pos_enc, pf, c->X, c->X32, c->XP, F, S, P + L, P, st);
  else
    nchw_to_tokens(in.vid, (long long)256 * P, c->X, c->X32, pos, pos_fs, c->XP, F, S, P + L, P, st);
  f.count(2);
  const size_t BL = (size_t)f.B * L;
  if (in.text_ids != nullptr) {
    VG_CUDA(cudaMemcpyAsync(c->ids_in, in.text_ids, BL * 4, cudaMemcpyDeviceToDevice, st));
    if (in.text_mask) VG_CUDA(cudaMemcpyAsync(c->tmask_in, in.text_mask, BL, cudaMemcpyDeviceToDevice, st));
  } else if (in.text_raw != nullptr) {
    f32_to_bf16(in.text_raw, c->traw, BL * c->ip_text.K, st);
    f.count();
  } else {
    VG_CUDA(cudaMemcpyAsync(c->text_in, in.text, BL * 256 * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (have_mask) { build_encoded_mask(in.vis_mask, in.text_mask, c->encmask, F, f.T, P, L, st); f.count(); }
}

// The captured part of the encoder phase: text tower / resizer (when the call brought raw text), text tokens, the encoder layers,
// the final norm and the pooled means.  Reads context-owned memory only (see ingest_encoder_inputs).
static void run_encoder(Fwd& f, int tkind, bool text_pad, bool have_mask, int pos_rows) {
  vgqa_ctx* c = f.c;
  cudaStream_t st = f.st;
  const int S = f.S, P = f.P, L = f.L, R = f.R, F = f.F;
  const float* text = c->text_in;
  if (tkind == 2) run_text_tower(f, c->ids_in, text_pad ? c->tmask_in : nullptr);   // → c->traw (bf16 last_hidden_state rows)
  if (tkind >= 1) {  // FeatureResizer: LayerNorm_1e-12(fc(hidden states)) (bert.py:90-96)
    GemmEpi ep; ep.C = c->tproj; ep.ldc = 256; ep.bias = c->ip_text.b; ep.bias_ld = 256; ep.C32 = c->tproj32; ep.ldc32 = 256;
    ep.ln_w = c->ip_text_ln.w; ep.ln_b = c->ip_text_ln.b; ep.ln_eps = 1e-12f;
    f.gemm(c->traw, c->ip_text.K, c->ip_text, f.B * L, ep);
    text = c->tproj32;
  }
  text_to_tokens(text, c->X, c->X32, c->XP, F, f.T, S, P, L, st);
  f.count();
  const uint8_t* km = have_mask ? c->encmask : nullptr;
  for (size_t l = 0; l < c->enc.size(); ++l) {
    EncLayer& e = c->enc[l];
    {  // q,k = (x + pos) Wqk^T + b ; v = x Wv^T + b   (modal_encoder.py:171-172) — one launch, two A operands
      GemmEpi ep; ep.C = c->QKV; ep.ldc = 768; ep.bias = e.qkv.b; ep.bias_ld = 768;
      gemm_ws(c->XP, c->X, 512, 256, e.qkv.W, 256, R, 768, 256, ep, st);
      f.count();
    }
    if (c->use_attn_tc && enc_attn_tc_supported(S))   // tcgen05/TMEM kernels: attn_tc.cu (S <= 128), attn_tc_long.cu (key tiles, online softmax)
      enc_attn_tc(c->QKV, c->AO, F, S, km, 0.17677669529663687f, st);
    else
      mha32(c->QKV, 768, c->QKV + 256, 768, c->QKV + 512, 768, c->AO, 256, F, S, S, km, 0.17677669529663687f, st);
    f.count();
    f.linear_res_ln(c->AO, 256, e.out, R, c->X32, e.ln1, 1e-5f, c->X1, 256, c->X1_32);
    // x = LN2(x1 + W2 relu(W1 x1 + b1) + b2); also emits x + pos for the next layer's Q/K projection
    const bool last = l + 1 == c->enc.size();
    if (c->use_ffn_fused && ffn_fused_supported(e.ff1.N)) {
      // one CTA-pair kernel, the [R, FFN_DIM] hidden activation never leaves the SMs (ffn_fused.cu)
      ffn_fused(c->X1, e.ff1.W, e.ff1.b, e.ff2.W, e.ff2.b, R, e.ff1.N, c->X1_32, e.ln2.w, e.ln2.b, 1e-5f, c->X, c->X32,
                last ? nullptr : c->XP, c->pos_enc, pos_rows, 0, st);
      f.count();
    } else {
      f.linear(c->X1, 256, e.ff1, R, c->HID, e.ff1.N, ACT_RELU);
      GemmEpi ep; ep.C = c->X; ep.ldc = 256; ep.bias = e.ff2.b; ep.bias_ld = 256; ep.res32 = c->X1_32; ep.ldres32 = 256;
      ep.C32 = c->X32; ep.ldc32 = 256; ep.ln_w = e.ln2.w; ep.ln_b = e.ln2.b; ep.ln_eps = 1e-5f;
      if (!last) { ep.C2 = c->XP; ep.ldc2 = 256; ep.add2 = c->pos_enc; ep.add2_period = pos_rows; }
      f.gemm(c->HID, e.ff2.K, e.ff2, R, ep);
    }
  }
  enc_finalize(c->X32, c->enc_norm.w, c->enc_norm.b, 1e-5f, c->Xf, c->frames_cls, c->pool[1], c->pool[0], c->pool32[1],
               c->pool32[0], F, S, P, L, st);
  text_sum(c->Xf, c->text_sums, f.B, f.T, S, P, L, st);
  f.all_reduce_f32(c->text_sums, (size_t)f.B * L * 256);
  text_finish(c->text_sums, 1.f / (float)f.T_global(), c->ftext, c->q0, c->q0_32, f.B, f.T, L, st);
  f.count(3);
}

// TemporalSampling (classifier.py:32-37) of classifier k (0 → t_temporal_clas on vid tokens, 1 → s_temporal_clas on vis tokens),
// enqueued on f.st
static void temporal_chain(Fwd& f, int k) {
  vgqa_ctx* c = f.c;
  const int F = f.F;
  const bf16* h = c->pool[k];
  const float* h32 = c->pool32[k];
  for (int i = 0; i < 2; ++i) {
    TsLayer& t = c->ts[k][i];
    f.linear(h, 256, t.q, F, c->c_q[k], 256);
    mha32(c->c_q[k], 256, c->kv_ts + t.kv_off, 2048, c->kv_ts + t.kv_off + 256, 2048, c->c_ctx[k], 256, f.B, f.T, f.L,
          nullptr, 0.17677669529663687f, f.st);
    f.count();
    f.linear_res_ln(c->c_ctx[k], 256, t.o, F, h32, t.ln_a, 1e-12f, c->c_a[k], 256, c->c_a32[k]);
    f.linear(c->c_a[k], 256, t.inter, F, c->c_i[k], 256, ACT_GELU);
    f.linear_res_ln(c->c_i[k], 256, t.outp, F, c->c_a32[k], t.ln_o, 1e-12f, c->c_h[k], 256, c->c_h32[k]);
    h = c->c_h[k]; h32 = c->c_h32[k];
  }
  Head& hd = c->ts_head[k];
  f.linear_res_ln(h, 256, hd.t, F, nullptr, hd.ln, 1e-12f, c->c_a[k], 256, nullptr, ACT_GELU);
  rowvec_head(c->c_a[k], 256, hd.dw, hd.db, c->logit_f[k], 1, F, 1, 0, f.st);
  f.count();
}

// The per-frame part of SpatialActivation k (classifier.py:64-81: two BertLayer_Cross blocks, head, attention map) on ALL frames,
// enqueued on f.st (scratch set k + 2).  It does not depend on which frames get chosen — every frame attends only to its own
// tokens — so it runs BESIDE TemporalSampling, and the second decoder pass reuses its logit rows / attention maps.
static void spatial_frames(Fwd& f, int k) {
  vgqa_ctx* c = f.c;
  const int F = f.F, P = f.P, S = f.S, z = k + 2;
  const int tok0 = k == 0 ? P + f.L : 0;  // t_* reads vid tokens, s_* reads vis tokens
  const bf16* q = c->q0;
  const float* q32 = c->q0_32;
  for (int i = 0; i < 2; ++i) {
    SaLayer& s = c->sa[k][i];
    f.linear(q, 256, s.qabs, F, c->c_qabs[k], 2048);
    xattn1(c->c_qabs[k], c->Xf + (size_t)tok0 * 256, S, F, P, nullptr, 0, nullptr, nullptr, 0, 0, nullptr, 0,
           0.17677669529663687f, c->c_ctx8[k], i == 1 ? c->attmap[k] : nullptr, f.st);
    f.count();
    f.linear_res_ln(c->c_ctx8[k], 2048, s.vo, F, q32, s.ln_a, 1e-12f, c->c_a[z], 256, c->c_a32[z]);
    f.linear(c->c_a[z], 256, s.inter, F, c->c_i[z], 256, ACT_GELU);
    f.linear_res_ln(c->c_i[z], 256, s.outp, F, c->c_a32[z], s.ln_o, 1e-12f, c->c_h[z], 256, c->c_h32[z]);
    q = c->c_h[z]; q32 = c->c_h32[z];
  }
  Head& hd = c->sa_head[k];
  f.linear_res_ln(q, 256, hd.t, F, nullptr, hd.ln, 1e-12f, c->c_a[z], 256, nullptr, ACT_GELU);
  rowvec_head(c->c_a[z], 256, hd.dw, hd.db, c->logit_rows[k], 64, F, hd.vocab, 0, f.st);
  f.count();
}

// The four classifier chains as four concurrent branches (graph branches under capture): TemporalSampling x2 ‖ SpatialActivation x2
static void run_classifiers(Fwd& f) {
  vgqa_ctx* c = f.c;
  f.linear(c->ftext, 256, c->ts_kv, f.B * f.L, c->kv_ts, 2048);
  cudaStream_t br[4] = {f.main, f.aux, f.aux2, f.aux3};
  for (int i = 1; i < 4; ++i) f.branch_fork(f.main, br[i]);
  for (int i = 0; i < 4; ++i) {
    f.st = br[i];
    if (i < 2) temporal_chain(f, i); else spatial_frames(f, i - 2);
  }
  for (int i = 1; i < 4; ++i) f.branch_join(br[i], f.main);
  f.st = f.main;
}

// Query seeding (grounding_net.py:131-136,155-160): masked means over the chosen frames of the per-frame classifier logits and of
// the attention-weighted tokens; both decoder passes.
static void run_spatial_seed(Fwd& f, const float* w, const float* K) {
  vgqa_ctx* c = f.c;
  const int F = f.F, P = f.P, S = f.S;
  f.fork();
  for (int k = 0; k < 2; ++k) {
    f.st = k == 0 ? f.main : f.aux;
    const int tok0 = k == 0 ? P + f.L : 0;
    Head& hd = c->sa_head[k];
    seed_partial(c->Xf, c->attmap[k], w, c->part[k], F, S, tok0, P, f.st);
    masked_sums(c->logit_rows[k], 64, hd.vocab, c->part[k], w, c->red[k], f.B, f.T, f.st);
    f.all_reduce_f32(c->red[k], (size_t)f.B * 320);
    // k = 0: init temporal query → TimeDecoder tgt; k = 1: init spatial query → PosDecoder tgt (cat cols 0..255)
    seed_finish(c->red[k], K, c->logits_r[k], hd.vocab, c->seedq[k], k == 0 ? c->t_tgt : c->p_cat, k == 0 ? 256 : 768,
                k == 0 ? c->t_tgt32 : c->p_tgt32, f.B, f.T, P, f.st);
    f.count(3);
  }
  f.join();
}

// second_pass: the anchors from frames_cls, the padded positional table, ca_kpos_proj(pos) and layer 0's query_pos are the
// same as in the first pass (they do not depend on the seeded queries) and are not recomputed.
static void run_decoders(Fwd& f, bool have_mask, int pos_frames, bool second_pass) {
  vgqa_ctx* c = f.c;
  cudaStream_t st = f.st;
  const int F = f.F, P = f.P, L = f.L, S = f.S, T = f.T, M = P + L;
  const int D = (int)c->tl.size();
  const long long pos_fs = pos_frames > 1 ? (long long)S * 256 : 0;
  // fused row-tile GEMM chains (chain.cu) for the frame-local tails of the decoder layers; a frame-sharded clip keeps the
  // one-launch-per-Linear form (its in-projection table is offset by the rank's first frame)
  const bool chain = c->use_chain && !f.sharded();
  const bool chain_head = c->use_chain_head && !chain;
  // anchors from frames_cls (query_decoder.py:92-94)
  if (!second_pass) {
    pos_fc_boxes(c->frames_cls, c->pfc_ln0w, c->pfc_ln0b, c->pfc_W, c->pfc_b, c->pfc_ln4w, c->pfc_ln4b, c->boxes0, F, st);
    f.count();
  }
  f.fork();
  // With a frame-invariant positional table the positional score terms are plain GEMMs over the absorbed queries:
  //   TimeDecoder  q~_h·(mem_m + pos_m) = q~_h·mem_m + (q~ pos^T)[h, m]
  //   PosDecoder   q2_h·kpos_h(m)       = (q2 · blockdiag(kpos)^T)[h, m]
  // so the cross-attention kernel streams the memory tokens once and only adds a [8, M] fp32 table per frame.
  const bool pos_gemm = pos_frames == 1;
  const int Npad = (M + 63) / 64 * 64, Mpad = (M + 7) / 8 * 8;
  if (pos_gemm && !second_pass) { pad_rows_bf16(c->pos_enc + (size_t)P * 256, c->pos_pad, M, Npad, st); f.count(); }
  // ---------------- TimeDecoder (query_decoder.py:379-486), memory = [text | vid] tokens
  for (int l = 0; l < D; ++l) {
    TimeLayer& t = c->tl[l];
    if (!(chain && l > 0)) {   // (the fused tail of layer l-1 already produced this layer's in-projection)
      GemmEpi ep; ep.C = c->t_qkv; ep.ldc = 768; ep.bias = t.tab + (size_t)c->sh_rank * T * 768; ep.bias_period = T; ep.bias_ld = 768;
      f.gemm(c->t_tgt, 256, t.qkv, F, ep);
    }
    {  // temporal self-attention across ALL frames of the clip: a sharded clip all-gathers the K|V rows (§8e)
      const bf16* kv = c->t_qkv;
      if (f.sharded()) { f.all_gather_bf16(c->t_qkv, c->t_qkv_all, (size_t)T * 768); kv = c->t_qkv_all; }
      mha32(c->t_qkv, 768, kv + 256, 768, kv + 512, 768, c->t_ao, 256, f.B, T, f.T_global(), nullptr, 0.17677669529663687f, st);
    }
    f.linear_res_ln(c->t_ao, 256, t.out, F, c->t_tgt32, t.ln1, 1e-5f, c->t_x, 256, c->t_x32);
    f.linear(c->t_x, 256, t.qabs, F, c->t_qabs, 2048);
    // keys = mem + pos_t (:474); mask = encoded_mask[:, :-P] applied positionally (:100,476)
    if (pos_gemm) {
      GemmEpi ep; ep.C = c->t_sb; ep.ldc = Npad; ep.c_f32 = 1;
      gemm_bf16_tn(c->t_qabs, 256, c->pos_pad, 256, F * 8, Npad, 256, ep, st);
      f.count();
      xattn1(c->t_qabs, c->Xf + (size_t)P * 256, S, F, M, nullptr, 0, nullptr, nullptr, 0, 0,
             have_mask ? c->encmask : nullptr, S, 0.17677669529663687f, c->t_ctx8, nullptr, st, c->t_sb, Npad);
    } else {
      xattn1(c->t_qabs, c->Xf + (size_t)P * 256, S, F, M, c->pos_enc + (size_t)P * 256, pos_fs, nullptr, nullptr, 0, 0,
             have_mask ? c->encmask : nullptr, S, 0.17677669529663687f, c->t_ctx8, nullptr, st);
    }
    if (chain) {
      // ONE launch (chain.cu): [vo + LN3] → [linear1 + ReLU → linear2 + LN4] → time_decoder.norm → next layer's in-projection
      ChainParams cp;
      cp.M = F; cp.T = T;
      chain_set_tmap(cp, 0, c->t_ctx8, F, 2048, 2048);
      ChOp& o0 = cp.ops[0];
      o0.a_kind = CH_A_STREAM; o0.a_tm = 0; o0.w = t.vo_t; o0.nkb = 32; o0.nblk = 2; o0.cn = 2; o0.acc0 = 0; o0.bias = t.vo.b;
      o0.res32 = c->t_x32; o0.ln_w = t.ln3.w; o0.ln_b = t.ln3.b; o0.out32 = c->t_x2_32; o0.out_act = 0;
      ChOp& o1 = cp.ops[1];
      o1.kind = CH_FFN; o1.a_blk = 0; o1.a_wait = CH_WAIT_EPI; o1.w = t.ff1_t; o1.bias = t.ff1.b; o1.w2 = t.ff2_t; o1.bias2 = t.ff2.b;
      o1.ff_chunks = t.ff1.N / 128; o1.res32 = c->t_x2_32; o1.ln_w = t.ln4.w; o1.ln_b = t.ln4.b; o1.out32 = c->t_tgt32; o1.out_act = 0;
      o1.ln2_w = c->time_norm.w; o1.ln2_b = c->time_norm.b; o1.out2 = c->t_inter + (size_t)l * F * 256; o1.ld_out2 = 256;   // :412
      cp.n_ops = 2;
      if (l + 1 < D) {
        TimeLayer& tn = c->tl[l + 1];
        ChOp& o2 = cp.ops[2];
        o2.a_blk = 0; o2.a_wait = CH_WAIT_EPI; o2.w = tn.qkv_t; o2.nkb = 4; o2.nblk = 6; o2.cn = 2; o2.acc0 = 1;
        o2.table = tn.tab; o2.table_ld = 768; o2.out_bf16 = c->t_qkv; o2.ld_out = 768;
        cp.n_ops = 3;
      }
      chain_launch(cp, st);
      f.count();
    } else {
      f.linear_res_ln(c->t_ctx8, 2048, t.vo, F, c->t_x32, t.ln3, 1e-5f, c->t_x2, 256, c->t_x2_32);
      f.linear(c->t_x2, 256, t.ff1, F, c->t_hid, t.ff1.N, ACT_RELU);
      f.linear_res_ln(c->t_hid, t.ff2.K, t.ff2, F, c->t_x2_32, t.ln4, 1e-5f, c->t_tgt, 256, c->t_tgt32);
      ln_rows(c->t_tgt32, 256, c->time_norm.w, c->time_norm.b, 1e-5f, c->t_inter + (size_t)l * F * 256, 256, F, st);  // :412
      f.count(3);
    }
  }
  // ---------------- PosDecoder (query_decoder.py:129-375), memory = [vis | text] tokens — second branch
  f.st = f.aux;
  st = f.aux;
  if (!second_pass) {
    GemmEpi ep; ep.C = c->kposb; ep.ldc = 1536; ep.bias = c->kpos_all.b; ep.bias_ld = c->kpos_all.N;
    f.gemm(c->pos_enc, 256, c->kpos_all, pos_frames * S, ep);   // ca_kpos_proj(pos_s) for all layers (:309)
    if (pos_gemm) { build_kpos_blockdiag(c->kposb, 1536, c->kpos_bd, D, M, Mpad, st); f.count(); }
  }
  const float* boxes = c->boxes0;
  for (int l = 0; l < D; ++l) {
    PosLayer& q = c->pl[l];
    // query_scale(tgt) (:176-179) only needs the layer input: it runs on a side stream beside the sine / ref_point_head /
    // self-attention chain and is joined where its product (the scaled sine embedding) is first used.
    const bf16* s256 = c->p_sine;
    int lds = 512;
    const bool have_qpos = chain_head && l > 0;   // the head chain of layer l-1 produced p_sine and query_pos already
    if (!have_qpos) { sine_embed(boxes, c->p_sine, F, st); f.count(); }         // :169
    if (l > 0) {
      f.side_fork(st);
      f.st = f.aux2;
      f.linear(c->p_cat, 768, c->qs0, F, c->p_h2, 256, ACT_RELU);
      GemmEpi ep; ep.C = c->p_s256; ep.ldc = 256; ep.bias = c->qs1.b; ep.bias_ld = 256; ep.mul = c->p_sine; ep.ldmul = 512;
      f.gemm(c->p_h2, 256, c->qs1, F, ep);
      f.st = st;
      s256 = c->p_s256; lds = 256;
    }
    if (!have_qpos) {
      f.linear(c->p_sine, 512, c->rph0, F, c->p_h, 256, ACT_RELU);              // ref_point_head (:170)
      f.linear(c->p_h, 256, c->rph1, F, c->p_cat + 256, 768);
    }
    { GemmEpi ep; ep.C = c->p_qkv; ep.ldc = 768; ep.bias = q.tab_sa + (size_t)c->sh_rank * T * 768; ep.bias_period = T; ep.bias_ld = 768;
      f.gemm(c->p_cat, 768, q.sa, F, ep); }                                     // 7 sa_* projs ∘ in_proj (:282-294)
    {
      const bf16* kv = c->p_qkv;
      if (f.sharded()) { f.all_gather_bf16(c->p_qkv, c->p_qkv_all, (size_t)T * 768); kv = c->p_qkv_all; }
      mha32(c->p_qkv, 768, kv + 256, 768, kv + 512, 768, c->p_ao, 256, f.B, T, f.T_global(), nullptr, 0.17677669529663687f, st);
    }
    f.linear_res_ln(c->p_ao, 256, q.sa_out, F, c->p_tgt32, q.ln1, 1e-5f, c->p_cat + 512, 768, c->p_x32);  // x → cat[:,512:]
    if (l > 0) f.side_join(st);
    const bf16* qa = l == 0 ? c->p_cat + 256 : c->p_cat + 512;                  // [qpos | x] or x
    // the absorbed content query (qabs) and the positional score term (sine proj → block-diagonal GEMM) are independent
    f.side_fork(st);
    f.st = f.aux2;
    f.linear(qa, 768, q.qabs, F, c->p_qabs, 2048);
    f.st = st;
    const bf16* q2res = nullptr;
    if (l == 0) { f.linear(qa, 768, q.q, F, c->p_q, 256); q2res = c->p_q; }      // :305,311-313
    { GemmEpi ep; ep.C = c->p_q2; ep.ldc = 256; ep.bias = q.sine.b; ep.bias_ld = 256; ep.res = q2res; ep.ldres = 256;
      f.gemm(s256, lds, q.sine, F, ep); }                                       // ca_qpos_sine_proj (:320)
    if (pos_gemm) {
      GemmEpi ep; ep.C = c->p_sb; ep.ldc = 8 * Mpad; ep.c_f32 = 1;
      gemm_bf16_tn(c->p_q2, 256, c->kpos_bd + (size_t)l * 8 * Mpad * 256, 256, F, 8 * Mpad, 256, ep, st);
      f.count();
      f.side_join(st);
      xattn1(c->p_qabs, c->Xf, S, F, M, nullptr, 0, nullptr, nullptr, 0, 0, nullptr, 0, 0.125f, c->p_ctx8, nullptr, st,
             c->p_sb, Mpad);                                                    // (512/8)^-0.5 (attention.py:151)
    } else {
      f.side_join(st);
      xattn1(c->p_qabs, c->Xf, S, F, M, nullptr, 0, c->p_q2, c->kposb + (size_t)l * 256, 1536, (long long)S * 1536,
             nullptr, 0, 0.125f, c->p_ctx8, nullptr, st);
    }
    f.count();
    float* anc = c->anchors + (size_t)l * F * 4;
    if (chain) {
      // ONE launch (chain.cu): [vo + LN3] → [linear1 + ReLU → linear2 + LN4] → bbox_embed (256 → 256 → 256 → 4) → sigmoid
      ChainParams cp;
      cp.M = F; cp.T = T;
      chain_set_tmap(cp, 0, c->p_ctx8, F, 2048, 2048);
      ChOp& o0 = cp.ops[0];
      o0.a_kind = CH_A_STREAM; o0.a_tm = 0; o0.w = q.vo_t; o0.nkb = 32; o0.nblk = 2; o0.cn = 2; o0.acc0 = 0; o0.bias = q.vo.b;
      o0.res32 = c->p_x32; o0.ln_w = q.ln3.w; o0.ln_b = q.ln3.b; o0.out32 = c->p_x2_32; o0.out_act = 0;
      ChOp& o1 = cp.ops[1];
      o1.kind = CH_FFN; o1.a_blk = 0; o1.a_wait = CH_WAIT_EPI; o1.w = q.ff1_t; o1.bias = q.ff1.b; o1.w2 = q.ff2_t; o1.bias2 = q.ff2.b;
      o1.ff_chunks = q.ff1.N / 128; o1.res32 = c->p_x2_32; o1.ln_w = q.ln4.w; o1.ln_b = q.ln4.b; o1.out32 = c->p_tgt32; o1.out_act = 0;
      o1.out_bf16 = c->p_cat; o1.ld_out = 768;
      ChOp& o2 = cp.ops[2];                                                     // bbox_embed (:188-192)
      o2.a_blk = 0; o2.a_wait = CH_WAIT_EPI; o2.w = c->bb0_t; o2.nkb = 4; o2.nblk = 2; o2.cn = 2; o2.acc0 = 1; o2.bias = c->bb0.b;
      o2.act = ACT_RELU; o2.out_act = 4;
      ChOp& o3 = cp.ops[3];
      o3.a_blk = 4; o3.a_wait = CH_WAIT_EPI; o3.w = c->bb1_t; o3.nkb = 4; o3.nblk = 2; o3.cn = 2; o3.acc0 = 0; o3.bias = c->bb1.b;
      o3.act = ACT_RELU; o3.head_w = c->bb2w; o3.head_b = c->bb2b; o3.head_n = 4; o3.head_act = 1; o3.head_out = anc;
      cp.n_ops = 4;
      chain_launch(cp, st);
      f.count();
    } else {
      f.linear_res_ln(c->p_ctx8, 2048, q.vo, F, c->p_x32, q.ln3, 1e-5f, c->p_x2, 256, c->p_x2_32);
      f.linear(c->p_x2, 256, q.ff1, F, c->p_hid, q.ff1.N, ACT_RELU);
      f.linear_res_ln(c->p_hid, q.ff2.K, q.ff2, F, c->p_x2_32, q.ln4, 1e-5f, c->p_cat, 768, c->p_tgt32);
      if (chain_head) {
        // ONE launch (chain.cu): bbox_embed 256 → 256 → 256 → 4, sigmoid (:188-192), and — for the next layer — the box sine
        // embedding (:169) and ref_point_head 512 → 256 → 256 (:170): seven launches of the per-Linear form
        ChainParams cp;
        cp.M = F; cp.T = T;
        chain_set_tmap(cp, 0, c->p_cat, F, 256, 768);
        ChOp& o0 = cp.ops[0];
        o0.kind = CH_LOAD; o0.a_tm = 0; o0.a_col0 = 0; o0.a_blk = 0; o0.nkb = 4;
        ChOp& o1 = cp.ops[1];
        o1.a_blk = 0; o1.a_wait = CH_WAIT_TMA; o1.w = c->bb0_t; o1.nkb = 4; o1.nblk = 2; o1.cn = 2; o1.acc0 = 0; o1.bias = c->bb0.b;
        o1.act = ACT_RELU; o1.out_act = 4;
        ChOp& o2 = cp.ops[2];
        o2.a_blk = 4; o2.a_wait = CH_WAIT_EPI; o2.w = c->bb1_t; o2.nkb = 4; o2.nblk = 2; o2.cn = 2; o2.acc0 = 1; o2.bias = c->bb1.b;
        o2.act = ACT_RELU; o2.head_w = c->bb2w; o2.head_b = c->bb2b; o2.head_n = 4; o2.head_act = 1; o2.head_out = anc;
        cp.n_ops = 3;
        if (l + 1 < D) {
          o2.gen_sine = 1; o2.sine_out = c->p_sine;
          ChOp& o3 = cp.ops[3];
          o3.a_blk = 0; o3.a_wait = CH_WAIT_EPI; o3.w = c->rph0_t; o3.nkb = 8; o3.nblk = 2; o3.cn = 2; o3.acc0 = 0; o3.bias = c->rph0.b;
          o3.act = ACT_RELU; o3.out_act = 0;
          ChOp& o4 = cp.ops[4];
          o4.a_blk = 0; o4.a_wait = CH_WAIT_EPI; o4.w = c->rph1_t; o4.nkb = 4; o4.nblk = 2; o4.cn = 2; o4.acc0 = 1; o4.bias = c->rph1.b;
          o4.out_bf16 = c->p_cat + 256; o4.ld_out = 768;
          cp.n_ops = 5;
        }
        chain_launch(cp, st);
        f.count();
      } else {
        f.linear(c->p_cat, 768, c->bb0, F, c->p_b1, 256, ACT_RELU);               // bbox_embed (:188-192)
        f.linear(c->p_b1, 256, c->bb1, F, c->p_b2, 256, ACT_RELU);
        rowvec_head(c->p_b2, 256, c->bb2w, c->bb2b, anc, 4, F, 4, 1, st);
        f.count(3);
      }
    }
    boxes = anc;
  }
  f.join();
}

static std::string shape_str(const vgqa_inputs& in) {
  return "clips=" + std::to_string(in.clips) + " T=" + std::to_string(in.T) + " H=" + std::to_string(in.H) +
         " W=" + std::to_string(in.W) + " L=" + std::to_string(in.L);
}

static void check_inputs(vgqa_ctx* c, const vgqa_inputs& in) {
  VG_CHECK(c->finalized, "vgqa_finalize_weights has not been called");
  VG_CHECK(in.clips >= 1 && in.T >= 2 && in.H >= 1 && in.W >= 1 && in.L >= 1, "bad shape: " + shape_str(in));
  VG_CHECK(in.clips <= c->cfg.max_clips && in.T <= c->cfg.max_frames && in.H * in.W <= c->cfg.max_hw &&
               in.L <= c->cfg.max_text,
           "shape exceeds the context capacity: " + shape_str(in));
  VG_CHECK(in.T <= c->cfg.max_video_len + 1,
           "T exceeds INPUT.MAX_VIDEO_LEN+1 rows of the time embedding (reference raises RuntimeError too)");
  VG_CHECK((in.vis || in.vis_raw) && (in.vid || in.vid_raw) && (in.text || in.text_raw || in.text_ids),
           "vis/vid/text (or their *_raw / text_ids forms) must be non-null");
  VG_CHECK(!in.text_ids || (!c->tt.empty() && c->ip_text.K > 0),
           "text_ids needs the 'text_encoder.body' (RoBERTa) and 'text_encoder.resizer' weights");
  VG_CHECK(!in.text_ids || in.L <= 64, "text_ids: a query has at most 64 tokens");
  VG_CHECK(in.raw_layout == 0 || in.raw_layout == 1, "raw_layout must be 0 (NCHW fp32) or 1 (channels-last bf16)");
  VG_CHECK(in.feat_layout == 0 || in.feat_layout == 1, "feat_layout must be 0 (NCHW fp32) or 1 (channels-last bf16)");
  VG_CHECK(!in.vis_raw || (c->ip_vis.K > 0 && in.vis_raw_ch == c->ip_vis.K),
           "vis_raw needs the 'input_proj' weights and vis_raw_ch equal to their input channels");
  VG_CHECK(!in.vid_raw || (c->ip_vid.K > 0 && in.vid_raw_ch == c->ip_vid.K),
           "vid_raw needs the 'input_proj2' weights and vid_raw_ch equal to their input channels");
  VG_CHECK(!in.text_raw || (c->ip_text.K > 0 && in.text_raw_ch == c->ip_text.K),
           "text_raw needs the 'text_encoder.resizer' weights and text_raw_ch equal to their input features");
  VG_CHECK(in.pos == nullptr || in.pos_frames == 1 || in.pos_frames == in.clips * in.T, "pos_frames must be 1 or clips*T");
  if (c->sh_world > 1) {
    VG_CHECK(in.clips == 1, "frame sharding handles one clip per call");
    VG_CHECK(in.T * c->sh_world <= c->cfg.max_video_len + 1, "sharded clip exceeds INPUT.MAX_VIDEO_LEN+1 frames");
    VG_CHECK(in.ori_sizes_hw == nullptr, "PostProcess of a sharded clip runs on the gathered outputs (vgqa_postprocess)");
  }
}

static void init_fwd(Fwd& f, vgqa_ctx* c, const vgqa_inputs& in, int phase, cudaStream_t st) {
  if (!c->aux_stream) {
    int prio_lo = 0, prio_hi = 0;
    VG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    VG_CUDA(cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, prio_hi));
    VG_CUDA(cudaStreamCreateWithPriority(&c->aux2_stream, cudaStreamNonBlocking, prio_hi));
    VG_CUDA(cudaStreamCreateWithPriority(&c->aux3_stream, cudaStreamNonBlocking, prio_hi));
    for (auto& e : c->fj) VG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  f.c = c; f.st = st; f.main = st; const bool one_stream = c->sh_world > 1 && !p2p_ready(c->p2p);   // NCCL-callback sharding: collectives stay in program order
  f.aux = one_stream ? st : c->aux_stream; f.aux2 = one_stream ? st : c->aux2_stream; f.aux3 = one_stream ? st : c->aux3_stream; f.B = in.clips; f.T = in.T; f.P = in.H * in.W; f.L = in.L; f.S = 2 * f.P + f.L;
  f.F = f.B * f.T; f.R = f.F * f.S; f.phase = phase;
}

// The part of a phase that reads context-owned memory only (what a CUDA graph captures).  `in` carries the shape, the flags and —
// for the decoder phase — the STAGED copies of ori_sizes_hw / force_choose1/2 (see forward_async).
// phase 0: CrossModalEncoder (+ final norm, pooled means); phase 1: everything after it.
static void forward_phase(vgqa_ctx* c, const vgqa_inputs& in, int phase, cudaStream_t st) {
  Fwd f;
  init_fwd(f, c, in, phase, st);
  const int F = f.F, D = (int)c->tl.size();
  const bool have_mask = in.vis_mask != nullptr || in.text_mask != nullptr;
  const int pos_rows = in.pos_frames * f.S;
  if (phase == 0) {
    // Leave `dec_sms` SMs out of the encoder's persistent grids: the decoder phase of the previous batch (other stream,
    // higher priority) is a chain of small latency-bound launches that then runs beside the encoder instead of between
    // its kernels.
    struct BudgetGuard { ~BudgetGuard() { set_sm_budget(0); } } guard;   // reset even when a launch throws
    set_sm_budget(c->enc_sm_budget);
    run_encoder(f, text_kind(in), in.text_mask != nullptr, have_mask, pos_rows);
    return;
  }
  if (in.stop_after_encoder) return;
  run_classifiers(f);
  select_pass1(c->logit_f[0], c->logit_f[1], 0.45f, in.force_choose1, c->att_seq, c->w1, c->K1, f.B, f.T, st);
  f.all_reduce_f32(c->K1, f.B);
  select_finish(c->w1, c->K1, f.B, f.T, f.T_global(), st);
  f.count(2);
  run_spatial_seed(f, c->w1, c->K1);
  run_decoders(f, have_mask, in.pos_frames, false);
  if (in.iteration_rate < 0) {  // grounding_net.py:143-163
    f.linear(c->t_inter + (size_t)(D - 1) * F * 256, 256, c->action_embed.l0, F, c->t_hs, 256, ACT_RELU);
    rowvec_head(c->t_hs, 256, c->action_embed.w1, c->action_embed.b1, c->act1, 1, F, 1, 1, st);  // sigmoid
    select_pass2(c->act1, in.force_choose2, c->w2, c->K2, f.B, f.T, st);
    f.all_reduce_f32(c->K2, f.B);
    select_finish(c->w2, c->K2, f.B, f.T, f.T_global(), st);
    f.count(3);
    run_spatial_seed(f, c->w2, c->K2);
    run_decoders(f, have_mask, in.pos_frames, true);
  }
  // heads over all decoder layers (grounding_net.py:177-181)
  f.linear(c->t_inter, 256, c->temp_embed.l0, D * F, c->t_hs, 256, ACT_RELU);
  rowvec_head(c->t_hs, 256, c->temp_embed.w1, c->temp_embed.b1, c->sted_all, 2, D * F, 2, 0, st);
  f.linear(c->t_inter, 256, c->action_embed.l0, D * F, c->t_hs, 256, ACT_RELU);
  rowvec_head(c->t_hs, 256, c->action_embed.w1, c->action_embed.b1, c->act_all, 1, D * F, 1, 0, st);
  f.count(2);
  const float* last_boxes = c->anchors + (size_t)(D - 1) * F * 4;
  const float* last_sted = c->sted_all + (size_t)(D - 1) * F * 2;
  if (in.ori_sizes_hw != nullptr) {
    postprocess(last_boxes, last_sted, in.ori_sizes_hw, c->boxes_px, c->sted_idx, f.B, f.T, st);
    f.count();
  }
}

// Copies of the results into the caller's buffers — outside the captured part (they depend on the caller's pointers).
static void emit_outputs(vgqa_ctx* c, const vgqa_inputs& in, const vgqa_outputs& out, cudaStream_t st) {
  const int F = in.clips * in.T, D = (int)c->tl.size(), B = in.clips;
  if (in.stop_after_encoder) {
    if (out.frames_cls) VG_CUDA(cudaMemcpyAsync(out.frames_cls, c->frames_cls, (size_t)F * 256 * 4, cudaMemcpyDeviceToDevice, st));
    return;
  }
  const float* wfinal = in.iteration_rate < 0 ? c->w2 : c->w1;
  const float* last_boxes = c->anchors + (size_t)(D - 1) * F * 4;
  const float* last_sted = c->sted_all + (size_t)(D - 1) * F * 2;
  auto cp = [&](void* dst, const void* src, size_t bytes) {
    if (dst != nullptr) VG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
  };
  cp(out.pred_boxes, last_boxes, (size_t)F * 4 * 4);
  cp(out.pred_sted, last_sted, (size_t)F * 2 * 4);
  cp(out.pred_actioness, c->act_all + (size_t)(D - 1) * F, (size_t)F * 4);
  cp(out.logits_f_m, c->logit_f[0], (size_t)F * 4);
  cp(out.logits_f_a, c->logit_f[1], (size_t)F * 4);
  cp(out.att_sequences, c->att_seq, (size_t)F * 4);
  cp(out.aux_boxes, c->anchors, (size_t)D * F * 4 * 4);
  cp(out.aux_sted, c->sted_all, (size_t)D * F * 2 * 4);
  cp(out.aux_actioness, c->act_all, (size_t)D * F * 4);
  cp(out.choose1, c->w1, (size_t)F * 4);
  cp(out.choose2, wfinal, (size_t)F * 4);
  cp(out.actioness_pass1, c->act1, (size_t)F * 4);
  if (in.ori_sizes_hw != nullptr) {
    cp(out.boxes_px, c->boxes_px, (size_t)F * 4 * 4);
    cp(out.sted_idx, c->sted_idx, (size_t)B * 2 * 4);
  }
  if (out.logits_r_m) VG_CUDA(cudaMemcpy2DAsync(out.logits_r_m, c->cfg.mot_num * 4, c->logits_r[0], c->cfg.mot_num * 4,
                                                c->cfg.mot_num * 4, B, cudaMemcpyDeviceToDevice, st));
  if (out.logits_r_a) VG_CUDA(cudaMemcpy2DAsync(out.logits_r_a, c->cfg.app_num * 4, c->logits_r[1], c->cfg.app_num * 4,
                                                c->cfg.app_num * 4, B, cudaMemcpyDeviceToDevice, st));
  cp(out.frames_cls, c->frames_cls, (size_t)F * 256 * 4);
}

}  // namespace vg

// bf16 → fp32 debug copy of the encoder output
__global__ void bf16_to_f32_kernel(const vg::bf16* in, float* out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

extern "C" {

int vgqa_create(const vgqa_config* cfg, vgqa_ctx** out) {
  try {
    VG_CHECK(cfg != nullptr && out != nullptr, "null argument");
    VG_CHECK(cfg->hidden == 256 && cfg->heads == 8, "this build supports MODEL.VSTG.HIDDEN=256, HEADS=8 only");
    VG_CHECK(cfg->ffn_dim > 0 && cfg->ffn_dim % 256 == 0, "FFN_DIM must be a multiple of 256");
    VG_CHECK(cfg->enc_layers >= 1 && cfg->dec_layers >= 1, "layer counts must be >= 1");
    VG_CHECK(cfg->app_num >= 1 && cfg->app_num <= 64 && cfg->mot_num >= 1 && cfg->mot_num <= 64, "vocab sizes must be in [1,64]");
    VG_CHECK(cfg->max_clips >= 1 && cfg->max_frames >= 2 && cfg->max_hw >= 1 && cfg->max_text >= 1, "bad capacity");
    VG_CHECK(cfg->max_frames <= cfg->max_video_len + 1, "max_frames exceeds max_video_len + 1");
    int ndev = 0;
    VG_CUDA(cudaGetDeviceCount(&ndev));
    VG_CHECK(ndev > 0, "no CUDA device: vgqa_b200 has no CPU fallback");
    vgqa_ctx* c = new vgqa_ctx();
    c->cfg = *cfg;
    VG_CUDA(cudaGetDevice(&c->device));
    cudaDeviceProp prop;
    VG_CUDA(cudaGetDeviceProperties(&prop, c->device));
    VG_CHECK(prop.major == 10, std::string("vgqa_b200 is built for sm_100a only; found ") + prop.name);
    *out = c;
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

void vgqa_destroy(vgqa_ctx* c) {
  if (!c) return;
  for (auto& g : c->graphs) cudaGraphExecDestroy(g.second.exec);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->host_stream) cudaStreamDestroy(c->host_stream);
  if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
  for (auto& h : c->hs)
    if (h.done) cudaEventDestroy(h.done);
  if (c->enc_stream) cudaStreamDestroy(c->enc_stream);
  if (c->dec_stream) cudaStreamDestroy(c->dec_stream);
  for (auto& b : c->bd) {
    if (b.ev_in) cudaEventDestroy(b.ev_in);
    if (b.enc_done) cudaEventDestroy(b.enc_done);
    if (b.dec_done) cudaEventDestroy(b.dec_done);
  }
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->aux2_stream) cudaStreamDestroy(c->aux2_stream);
  if (c->aux3_stream) cudaStreamDestroy(c->aux3_stream);
  vg::p2p_destroy(c->p2p);
  c->swin.release();
  c->resnet.release();
  for (auto& e : c->fj) if (e) cudaEventDestroy(e);
  c->warena.release();
  c->ws.release();
  delete c;
}

int vgqa_set_weight(vgqa_ctx* c, const char* name, const float* data, const int64_t* shape, int ndim) {
  try {
    VG_CHECK(c && name && data && (shape || ndim == 0), "null argument");
    VG_CHECK(!c->finalized, "weights are already finalized");
    HostT t;
    t.shape.assign(shape, shape + ndim);
    t.v.assign(data, data + t.numel());
    c->sd[name] = std::move(t);
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_finalize_weights(vgqa_ctx* c) {
  try {
    VG_CHECK(c && !c->finalized, "bad context");
    vg::pack_weights(c);
    vg::alloc_workspace(c);
    c->sd.clear();
    c->finalized = true;
    VG_CUDA(cudaDeviceSynchronize());
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

}  // extern "C" (reopened below)

namespace vg {

static void ensure_streams(vgqa_ctx* c) {
  if (c->enc_stream) return;
  int prio_lo = 0, prio_hi = 0;
  VG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  // phase 1 is a long chain of small kernels: give it priority so that its blocks are placed first whenever SMs free up
  // (VGQA_DEC_PRIO: 1 = decoder high [default], 0 = equal, -1 = encoder high)
  int mode = 1;
  if (const char* e = getenv("VGQA_DEC_PRIO")) mode = atoi(e);
  VG_CUDA(cudaStreamCreateWithPriority(&c->enc_stream, cudaStreamNonBlocking, mode < 0 ? prio_hi : prio_lo));
  VG_CUDA(cudaStreamCreateWithPriority(&c->dec_stream, cudaStreamNonBlocking, mode > 0 ? prio_hi : prio_lo));
  VG_CUDA(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
  for (auto& b : c->bd) {
    VG_CUDA(cudaEventCreateWithFlags(&b.ev_in, cudaEventDisableTiming));
    VG_CUDA(cudaEventCreateWithFlags(&b.enc_done, cudaEventDisableTiming));
    VG_CUDA(cudaEventCreateWithFlags(&b.dec_done, cudaEventDisableTiming));
    if (c->timeline) for (cudaEvent_t* e : {&b.t_enc0, &b.t_enc1, &b.t_dec0, &b.t_dec1}) VG_CUDA(cudaEventCreate(e));
  }
  if (c->timeline) { VG_CUDA(cudaEventCreate(&c->t_base)); VG_CUDA(cudaEventRecord(c->t_base, c->enc_stream)); }
  for (auto& h : c->hs) VG_CUDA(cudaEventCreateWithFlags(&h.done, cudaEventDisableTiming));
}

// Enqueue the captured part of one phase on `ex`: eagerly, or as a cached CUDA graph.  The key holds the phase, the slot, the
// shape and the presence flags of the optional inputs — no pointers: everything a graph reads or writes is context-owned
// (ingest_encoder_inputs / the staging in forward_async / emit_outputs handle the caller's buffers outside of it).
static int run_phase(vgqa_ctx* c, const vgqa_inputs& in, int phase, int slot, cudaStream_t ex, bool eager) {
  c->launches = 0;
  if (eager) {
    forward_phase(c, in, phase, ex);
    return c->launches;
  }
  const std::vector<uint64_t> key = {
      (uint64_t)phase, (uint64_t)slot, (uint64_t)in.clips, (uint64_t)in.T, (uint64_t)in.H, (uint64_t)in.W, (uint64_t)in.L,
      (uint64_t)in.pos_frames, (uint64_t)(in.iteration_rate < 0), (uint64_t)text_kind(in), (uint64_t)(in.vis_mask != nullptr),
      (uint64_t)(in.text_mask != nullptr), (uint64_t)(in.ori_sizes_hw != nullptr), (uint64_t)(in.force_choose1 != nullptr),
      (uint64_t)(in.force_choose2 != nullptr), (uint64_t)c->sh_world, (uint64_t)c->sh_rank};
  auto it = c->graphs.find(key);
  if (it == c->graphs.end()) {
    // First call of this shape: run it eagerly — that IS this call's execution (it also sets the kernels' function attributes
    // and validates the launches) — and then capture the same sequence for the calls to come.  The captured part is executed
    // exactly once per call: the encoder layers update the ingested token rows in place, and a frame-sharded forward advances
    // the device-side exchange sequence on every execution, so a second run would neither be idempotent nor stay in step
    // with the other ranks.
    forward_phase(c, in, phase, ex);
    const int eager_launches = c->launches;
    VG_CUDA(cudaStreamSynchronize(ex));
    VG_CUDA(cudaStreamSynchronize(c->aux_stream));
    VG_CUDA(cudaStreamSynchronize(c->aux2_stream));
    VG_CUDA(cudaStreamSynchronize(c->aux3_stream));
    c->launches = 0;
    cudaGraph_t graph = nullptr;
    VG_CUDA(cudaStreamBeginCapture(ex, cudaStreamCaptureModeThreadLocal));
    try {
      forward_phase(c, in, phase, ex);
    } catch (...) {
      cudaStreamEndCapture(ex, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    VG_CUDA(cudaStreamEndCapture(ex, &graph));
    vgqa_ctx::GraphEntry ge;
    VG_CUDA(cudaGraphInstantiate(&ge.exec, graph, 0));
    cudaGraphDestroy(graph);
    ge.launches = c->launches;
    ++c->graph_captures;
    if (c->graphs.size() >= 64) {   // shapes only: a serving process sees a handful; bound the cache anyway
      VG_CUDA(cudaDeviceSynchronize());
      for (auto& g : c->graphs) cudaGraphExecDestroy(g.second.exec);
      c->graphs.clear();
    }
    c->graphs.emplace(key, ge);
    return eager_launches;
  }
  VG_CUDA(cudaGraphLaunch(it->second.exec, ex));
  return it->second.launches;
}

// Both phases of one call, for boundary slot `slot`, ordered after everything enqueued on `st` so far.
static void forward_async(vgqa_ctx* c, const vgqa_inputs& user_in, const vgqa_outputs& out, int slot, cudaStream_t st) {
  check_inputs(c, user_in);
  ensure_streams(c);
  select_slot(c, slot);
  vgqa_ctx::Boundary& b = c->bd[slot];
  vgqa_inputs in = user_in;
  // PositionEmbeddingSine generated in the library: one table for every frame unless a padding mask makes it per-frame
  if (in.pos == nullptr) in.pos_frames = in.vis_mask != nullptr ? in.clips * in.T : 1;
  // a sharded forward that exchanges through the host callback (NCCL) cannot be captured; the peer-memory exchange can
  const bool eager = !c->cfg.use_cuda_graph || out.encoded_feature != nullptr || in.stop_after_encoder ||
                     (c->sh_world > 1 && !p2p_ready(c->p2p));
  VG_CUDA(cudaEventRecord(b.ev_in, st));
  VG_CUDA(cudaStreamWaitEvent(c->enc_stream, b.ev_in, 0));
  if (b.used) VG_CUDA(cudaStreamWaitEvent(c->enc_stream, b.dec_done, 0));  // phase 1 of the previous user of this slot
  if (c->timeline) VG_CUDA(cudaEventRecord(b.t_enc0, c->enc_stream));
  int launches = 0;
  {  // ---- encoder phase: ingest the caller's tensors (eager), then the captured part
    Fwd f;
    init_fwd(f, c, in, 0, c->enc_stream);
    c->launches = 0;
    ingest_encoder_inputs(f, in, in.vis_mask != nullptr || in.text_mask != nullptr);
    launches += c->launches;
    launches += run_phase(c, in, 0, slot, c->enc_stream, eager);
  }
  if (c->timeline) VG_CUDA(cudaEventRecord(b.t_enc1, c->enc_stream));
  VG_CUDA(cudaEventRecord(b.enc_done, c->enc_stream));
  VG_CUDA(cudaStreamWaitEvent(c->dec_stream, b.enc_done, 0));
  if (c->timeline) VG_CUDA(cudaEventRecord(b.t_dec0, c->dec_stream));
  {  // ---- decoder phase: stage its (tiny) optional inputs, the captured part, then the copies into the caller's outputs
    vgqa_ctx::InStage& is = c->ins[slot];
    const size_t F = (size_t)in.clips * in.T;
    auto stage = [&](const float*& ptr, float* dst, size_t n) {
      if (ptr == nullptr) return;
      VG_CUDA(cudaMemcpyAsync(dst, ptr, n * 4, cudaMemcpyDeviceToDevice, c->dec_stream));
      ptr = dst;
    };
    stage(in.ori_sizes_hw, is.sizes, (size_t)in.clips * 2);
    stage(in.force_choose1, is.f1, F);
    stage(in.force_choose2, is.f2, F);
    launches += run_phase(c, in, 1, slot, c->dec_stream, eager);
    emit_outputs(c, in, out, c->dec_stream);
  }
  if (c->timeline) VG_CUDA(cudaEventRecord(b.t_dec1, c->dec_stream));
  if (out.encoded_feature) {
    const size_t n = (size_t)in.clips * in.T * (2 * in.H * in.W + in.L) * 256;
    bf16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->dec_stream>>>(c->Xf, out.encoded_feature, n);
  }
  VG_CUDA(cudaEventRecord(b.dec_done, c->dec_stream));
  b.used = true;
  c->last_launches = launches;
}

}  // namespace vg

extern "C" {

int vgqa_forward_async(vgqa_ctx* c, const vgqa_inputs* in, const vgqa_outputs* out, int slot, void* stream) {
  try {
    VG_CHECK(c && in && out && (slot == 0 || slot == 1), "bad argument");
    vg::forward_async(c, *in, *out, slot, static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_forward_wait(vgqa_ctx* c, int slot, void* stream, int host_sync) {
  try {
    VG_CHECK(c && (slot == 0 || slot == 1), "bad argument");
    if (!c->bd[slot].used) return 0;
    if (host_sync) VG_CUDA(cudaEventSynchronize(c->bd[slot].dec_done));
    else VG_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), c->bd[slot].dec_done, 0));
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_forward(vgqa_ctx* c, const vgqa_inputs* in, const vgqa_outputs* out, void* stream) {
  int rc = vgqa_forward_async(c, in, out, 0, stream);
  return rc != 0 ? rc : vgqa_forward_wait(c, 0, stream, 0);
}

int vgqa_forward_host_async(vgqa_ctx* c, const vgqa_inputs* hin, const vgqa_outputs* hout, int slot) {
  try {
    VG_CHECK(c && hin && hout && (slot == 0 || slot == 1), "bad argument");
    vg::check_inputs(c, *hin);
    VG_CHECK(hout->encoded_feature == nullptr, "encoded_feature is only available through vgqa_forward");
    vg::ensure_streams(c);
    vgqa_ctx::HostSlot& h = c->hs[slot];
    cudaStream_t up = c->h2d_stream;
    const size_t B = hin->clips, T = hin->T, P = (size_t)hin->H * hin->W, L = hin->L, F = B * T;
    const size_t D = c->tl.size();
    // The feature / text / mask staging buffers of this slot are read by phase 0 (run_encoder) only: their upload may start as
    // soon as the ENCODER phase of the previous call on this slot is done, i.e. it overlaps that call's decoder phase and the
    // other slot's encoder phase.  (Waiting for dec_done here delayed the next encoder phase by most of the 8 ms upload.)
    if (c->bd[slot].used) VG_CUDA(cudaStreamWaitEvent(up, c->bd[slot].enc_done, 0));
    auto h2d = [&](void* dst, const void* src, size_t bytes) {
      VG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, up));
    };
    vgqa_inputs din = *hin;
    const size_t raw_es = hin->raw_layout == 1 ? 2 : 4;   // channels-last bf16 or NCHW fp32 maps
    if (hin->vis_raw) { h2d(h.vis_raw, hin->vis_raw, F * c->ip_vis.K * P * raw_es); din.vis_raw = h.vis_raw; din.vis = nullptr; }
    else { h2d(h.vis, hin->vis, F * 256 * P * (hin->feat_layout == 1 ? 2 : 4)); din.vis = h.vis; }
    if (hin->vid_raw) { h2d(h.vid_raw, hin->vid_raw, F * c->ip_vid.K * P * raw_es); din.vid_raw = h.vid_raw; din.vid = nullptr; }
    else { h2d(h.vid, hin->vid, F * 256 * P * (hin->feat_layout == 1 ? 2 : 4)); din.vid = h.vid; }
    if (hin->text_ids) { h2d(h.ids, hin->text_ids, B * L * 4); din.text_ids = h.ids; din.text = nullptr; din.text_raw = nullptr; }
    else if (hin->text_raw) { h2d(h.text_raw, hin->text_raw, B * L * c->ip_text.K * 4); din.text_raw = h.text_raw; din.text = nullptr; }
    else { h2d(h.text, hin->text, B * L * 256 * 4); din.text = h.text; }
    if (hin->pos) { h2d(h.pos, hin->pos, (size_t)hin->pos_frames * 256 * P * 4); din.pos = h.pos; }
    if (hin->vis_mask) { h2d(h.vmask, hin->vis_mask, F * P); din.vis_mask = h.vmask; }
    if (hin->text_mask) { h2d(h.tmask, hin->text_mask, B * L); din.text_mask = h.tmask; }
    // sizes / forced selections are read by phase 1: wait for the previous call's decoder phase (and result downloads)
    if (c->bd[slot].used) VG_CUDA(cudaStreamWaitEvent(up, c->bd[slot].dec_done, 0));
    if (hin->ori_sizes_hw) { h2d(h.sizes, hin->ori_sizes_hw, B * 2 * 4); din.ori_sizes_hw = h.sizes; }
    if (hin->force_choose1) { h2d(h.f1, hin->force_choose1, F * 4); din.force_choose1 = h.f1; }
    if (hin->force_choose2) { h2d(h.f2, hin->force_choose2, F * 4); din.force_choose2 = h.f2; }
    vgqa_outputs none;
    std::memset(&none, 0, sizeof(none));
    vg::forward_async(c, din, none, slot, up);
    cudaStream_t st = c->dec_stream;   // results are downloaded right behind phase 1, before the next call's phase 1
    auto d2h = [&](void* dst, const void* src, size_t bytes) {
      if (dst != nullptr) VG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    };
    d2h(hout->pred_boxes, c->anchors + (D - 1) * F * 4, F * 16);
    d2h(hout->pred_sted, c->sted_all + (D - 1) * F * 2, F * 8);
    d2h(hout->pred_actioness, c->act_all + (D - 1) * F, F * 4);
    d2h(hout->logits_f_m, c->logit_f[0], F * 4);
    d2h(hout->logits_f_a, c->logit_f[1], F * 4);
    d2h(hout->logits_r_m, c->logits_r[0], B * c->cfg.mot_num * 4);
    d2h(hout->logits_r_a, c->logits_r[1], B * c->cfg.app_num * 4);
    d2h(hout->att_sequences, c->att_seq, F * 4);
    d2h(hout->aux_boxes, c->anchors, D * F * 16);
    d2h(hout->aux_sted, c->sted_all, D * F * 8);
    d2h(hout->aux_actioness, c->act_all, D * F * 4);
    d2h(hout->choose1, c->w1, F * 4);
    d2h(hout->choose2, hin->iteration_rate < 0 ? c->w2 : c->w1, F * 4);
    d2h(hout->actioness_pass1, c->act1, F * 4);
    if (hin->ori_sizes_hw) {
      d2h(hout->boxes_px, c->boxes_px, F * 16);
      d2h(hout->sted_idx, c->sted_idx, B * 8);
    }
    d2h(hout->frames_cls, c->frames_cls, F * 256 * 4);
    VG_CUDA(cudaEventRecord(h.done, st));
    VG_CUDA(cudaEventRecord(c->bd[slot].dec_done, st));   // the slot is busy until the downloads are done, too
    h.used = true;
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_forward_host_wait(vgqa_ctx* c, int slot) {
  try {
    VG_CHECK(c && (slot == 0 || slot == 1), "bad argument");
    if (c->hs[slot].used) VG_CUDA(cudaEventSynchronize(c->hs[slot].done));
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_forward_host(vgqa_ctx* c, const vgqa_inputs* hin, const vgqa_outputs* hout) {
  int rc = vgqa_forward_host_async(c, hin, hout, 0);
  return rc != 0 ? rc : vgqa_forward_host_wait(c, 0);
}

int vgqa_set_sharding(vgqa_ctx* c, int rank, int world, vgqa_exchange_fn fn, void* user) {
  try {
    VG_CHECK(c && world >= 1 && rank >= 0 && rank < world, "bad rank / world");
    VG_CHECK(world == 1 || fn != nullptr, "a sharded context needs an exchange callback");
    c->sh_rank = rank; c->sh_world = world; c->sh_fn = fn; c->sh_user = user;
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_shard_p2p_export(vgqa_ctx* c, int rank, int world, unsigned char* handle_out) {
  try {
    VG_CHECK(c && handle_out, "bad argument");
    const vgqa_config& g = c->cfg;
    long long slot = (long long)g.max_frames * 768 * 2;                                  // temporal-attention Q|K|V rows
    slot = std::max(slot, (long long)g.max_clips * g.max_text * 256 * 4);                // text-token sums
    slot = std::max(slot, (long long)g.max_clips * 320 * 4);                             // classifier / seed sums
    vg::p2p_export(c->p2p, rank, world, slot, handle_out);
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_shard_p2p_import(vgqa_ctx* c, int rank, int world, const unsigned char* handles) {
  try {
    VG_CHECK(c && handles && world >= 2 && rank >= 0 && rank < world, "bad argument");
    vg::p2p_import(c->p2p, handles);
    c->sh_rank = rank; c->sh_world = world;
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_shard_p2p_error(vgqa_ctx* c) {
  try { return c ? vg::p2p_error(c->p2p) : 0; } catch (const std::exception& e) { vg::set_last_error(e.what()); return -1; }
}

int vgqa_postprocess(const float* boxes, const float* sted, const float* sizes_hw, float* boxes_px, int32_t* sted_idx,
                     int clips, int T, void* stream) {
  try {
    VG_CHECK(boxes && sted && sizes_hw && boxes_px && sted_idx && clips >= 1 && T >= 2, "vgqa_postprocess: bad argument");
    vg::postprocess(boxes, sted, sizes_hw, boxes_px, sted_idx, clips, T, static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_text_tower_hidden(const vgqa_ctx* c) { return c ? c->tt_hd : 0; }

int vgqa_text_tower(vgqa_ctx* c, const int32_t* ids, const uint8_t* text_mask, int clips, int L, float* hidden, float* text,
                    void* stream) {
  try {
    VG_CHECK(c && c->finalized && ids && clips >= 1 && L >= 1, "vgqa_text_tower: bad argument");
    VG_CHECK(!c->tt.empty() && c->ip_text.K > 0, "vgqa_text_tower needs the 'text_encoder.body' and 'text_encoder.resizer' weights");
    VG_CHECK(clips <= c->cfg.max_clips && L <= c->cfg.max_text && L <= 64, "vgqa_text_tower: shape exceeds the context capacity");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    vg::Fwd f;
    f.c = c; f.st = st; f.main = st; f.aux = st; f.aux2 = st; f.aux3 = st; f.B = clips; f.L = L;
    vg::run_text_tower(f, ids, text_mask);
    vg::GemmEpi ep; ep.C = c->tproj; ep.ldc = 256; ep.bias = c->ip_text.b; ep.bias_ld = 256; ep.C32 = c->tproj32; ep.ldc32 = 256;
    ep.ln_w = c->ip_text_ln.w; ep.ln_b = c->ip_text_ln.b; ep.ln_eps = 1e-12f;
    f.gemm(c->traw, c->ip_text.K, c->ip_text, clips * L, ep);
    const size_t n = (size_t)clips * L * c->tt_hd;
    if (hidden) bf16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->traw, hidden, n);
    if (text) VG_CUDA(cudaMemcpyAsync(text, c->tproj32, (size_t)clips * L * 256 * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_swin_stage(vgqa_ctx* c, const float* x, int clips, int T, int H, int W, void* out_bf16, float* out_f32, void* stream) {
  try {
    VG_CHECK(c && c->finalized && x && (out_bf16 || out_f32), "vgqa_swin_stage: bad argument");
    c->last_launches = c->swin.forward_stage4(x, clips, T, H, W, static_cast<vg::bf16*>(out_bf16), out_f32, static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_swin_backbone(vgqa_ctx* c, const float* frames, int clips, int T, int R, void* out_bf16, float* out_f32, float* const* stage_out,
                       void* stream) {
  try {
    VG_CHECK(c && c->finalized && frames && (out_bf16 || out_f32), "vgqa_swin_backbone: bad argument");
    c->last_launches = c->swin.forward_full(frames, clips, T, R, static_cast<vg::bf16*>(out_bf16), out_f32, stage_out,
                                            static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_resnet_backbone(vgqa_ctx* c, const float* frames, int n_frames, int R, void* out_bf16, float* out_f32, float* const* layer_out,
                         void* stream) {
  try {
    VG_CHECK(c && c->finalized && frames && (out_bf16 || out_f32), "vgqa_resnet_backbone: bad argument");
    c->last_launches = c->resnet.forward(frames, n_frames, R, static_cast<vg::bf16*>(out_bf16), out_f32, layer_out,
                                         static_cast<cudaStream_t>(stream));
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

int vgqa_last_launch_count(const vgqa_ctx* c) { return c ? c->last_launches : 0; }

int vgqa_graph_capture_count(const vgqa_ctx* c) { return c ? c->graph_captures : 0; }

int vgqa_debug_phase_times(vgqa_ctx* c, int slot, float* ms4) {
  try {
    VG_CHECK(c && ms4 && (slot == 0 || slot == 1) && c->timeline && c->bd[slot].used, "phase timeline is off (VGQA_TIMELINE=1)");
    vgqa_ctx::Boundary& b = c->bd[slot];
    VG_CUDA(cudaEventSynchronize(b.t_dec1));
    cudaEvent_t ev[4] = {b.t_enc0, b.t_enc1, b.t_dec0, b.t_dec1};
    for (int i = 0; i < 4; ++i) VG_CUDA(cudaEventElapsedTime(&ms4[i], c->t_base, ev[i]));
    return 0;
  } catch (const std::exception& e) { vg::set_last_error(e.what()); return 1; }
}

double vgqa_reference_flops(int T, int H, int W, int L, int enc_layers, int dec_layers, int ffn, int passes) {
  // 2*MACs of the reference modules: encoder + TemporalSampling + passes*(SpatialActivation + decoders)
  const double d = 256, P = (double)H * W, S = 2 * P + L, M = P + L, F = ffn, Tt = T;
  const double enc = enc_layers * (Tt * S * (2 * d * 3 * d + 2 * d * d + 4 * d * F) + Tt * 4 * S * S * d);
  const double ts = 2 * (2 * (2 * (Tt * d * d + 2 * L * d * d) + 4 * Tt * L * d + 2 * 3 * Tt * d * d) + 2 * Tt * d * d + 2 * Tt * d);
  const double sp = 2 * (2 * (2 * (Tt * d * d + 2 * Tt * P * d * d) + 4 * Tt * P * d + 2 * 3 * Tt * d * d) + 2 * Tt * d * d);
  const double qside = 2 * Tt * d * d;  // one 256x256 Linear over T rows
  const double timed = dec_layers * (2 * 2 * Tt * M * d * d + 4 * qside + 4 * Tt * Tt * d + qside + 4 * Tt * M * d + qside +
                                     4 * Tt * d * F);
  const double posd = dec_layers * (3 * 2 * Tt * M * d * d + 11 * qside + 4 * Tt * Tt * d + 4 * Tt * M * d * 1.5 + 3 * qside +
                                    4 * Tt * d * F + 4 * qside + 2 * qside);
  return enc + ts + passes * (sp + timed + posd);
}

}  // extern "C"
