"""CPU ORACLE — test infrastructure only, NOT a product path.

A plain-numpy (fp32) restatement of the VGQA grounding hot path (everything `VSTGNet.forward` does
after the ResNet101 / Video-Swin / RoBERTa feature extractors, plus `PostProcess` and the tube/segment
merge).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this file, and only as the checker / the reported CPU baseline.  The product
(`vgqa_b200`) never imports it and fails loudly when its CUDA library is missing.

Pinning: the reference ships no tests / golden vectors, so the oracle is pinned against
outputs of the reference's own PyTorch modules imported in the build container
(`tests/golden/make_golden.py` → `tests/golden/*.npz`, checked by `tests/test_oracle_golden.py`; with
`/root/reference` present `tests/test_oracle_vs_reference.py` also compares live).

Every function cites the reference file:line it restates (paths relative to the reference root).
Weights are a dict {reference state_dict key: np.ndarray fp32} (VSTGNet key names).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32
try:  # vectorised erf for the BERT-style GELU; scipy is in the image
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float64])


# ----------------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------------
def linear(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray] = None) -> np.ndarray:
    """torch.nn.functional.linear: x @ w.T + b."""
    y = x @ w.T
    if b is not None:
        y = y + b
    return y.astype(F32, copy=False)


def layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """nn.LayerNorm (biased variance, eps inside sqrt) — vgqa/core/decoder/modal_encoder.py:154-155."""
    x = x.astype(F32, copy=False)
    u = x.mean(-1, keepdims=True)
    s = ((x - u) ** 2).mean(-1, keepdims=True)
    return ((x - u) / np.sqrt(s + F32(eps)) * w + b).astype(F32)


def bert_layer_norm(x, w, b, eps: float = 1e-12):
    """BertLayerNorm — vgqa/core/language/bert_module.py:18-31 (TF style, eps 1e-12 inside sqrt)."""
    return layer_norm(x, w, b, eps)


def gelu_erf(x: np.ndarray) -> np.ndarray:
    """vgqa/core/language/bert_module.py:13-15."""
    return (x * 0.5 * (1.0 + _erf(x / math.sqrt(2.0)))).astype(F32)


def sigmoid(x: np.ndarray) -> np.ndarray:
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(F32)


def softmax(x: np.ndarray, axis: int = -1) -> np.ndarray:
    m = x.max(axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    e = np.exp(x - m)
    return (e / e.sum(axis=axis, keepdims=True)).astype(F32)


def log_softmax(x: np.ndarray, axis: int = -1) -> np.ndarray:
    m = x.max(axis=axis, keepdims=True)
    y = x - m
    return (y - np.log(np.exp(y).sum(axis=axis, keepdims=True))).astype(F32)


def mlp(x: np.ndarray, sd: Dict[str, np.ndarray], prefix: str, num_layers: int) -> np.ndarray:
    """MLP — vgqa/core/model_utils.py:43-58 (ReLU between layers, dropout = identity in eval)."""
    for i in range(num_layers):
        x = linear(x, sd[f"{prefix}.layers.{i}.weight"], sd[f"{prefix}.layers.{i}.bias"])
        if i < num_layers - 1:
            x = np.maximum(x, 0)
    return x


def torch_mha(query, key, value, in_w, in_b, out_w, out_b, nhead: int,
              key_padding_mask: Optional[np.ndarray] = None) -> np.ndarray:
    """nn.MultiheadAttention forward, seq-first (len, batch, E), packed in_proj, eval mode.

    Used at vgqa/core/decoder/modal_encoder.py:148,172 and query_decoder.py:224,294,434-435,469,472.
    q is scaled by head_dim**-0.5; bool key_padding_mask (batch, src) → -inf; returns attn output only.
    """
    lq, bsz, e = query.shape
    lk = key.shape[0]
    dh = e // nhead
    q = linear(query, in_w[:e], in_b[:e])
    k = linear(key, in_w[e:2 * e], in_b[e:2 * e])
    v = linear(value, in_w[2 * e:], in_b[2 * e:])
    q = q.reshape(lq, bsz * nhead, dh).transpose(1, 0, 2) * F32(dh ** -0.5)
    k = k.reshape(lk, bsz * nhead, dh).transpose(1, 0, 2)
    v = v.reshape(lk, bsz * nhead, dh).transpose(1, 0, 2)
    s = q @ k.transpose(0, 2, 1)  # (bsz*nhead, lq, lk)
    if key_padding_mask is not None:
        s = s.reshape(bsz, nhead, lq, lk)
        s = np.where(key_padding_mask[:, None, None, :], -np.inf, s)
        s = s.reshape(bsz * nhead, lq, lk)
    p = softmax(s, -1)
    o = (p @ v).transpose(1, 0, 2).reshape(lq, bsz, e)
    return linear(o, out_w, out_b)


# ----------------------------------------------------------------------------------------------
# constants: position tables
# ----------------------------------------------------------------------------------------------
def position_embedding_sine(mask: np.ndarray, num_pos_feats: int = 128, temperature: float = 10000.0) -> np.ndarray:
    """PositionEmbeddingSine(128, normalize=True) — vgqa/core/vision/position_encoding.py:50-91,131-136.

    mask: bool (T,H,W), True = padded. Returns (T, 256, H, W) fp32."""
    not_mask = ~mask
    y_embed = not_mask.cumsum(1).astype(F32)
    x_embed = not_mask.cumsum(2).astype(F32)
    eps, scale = F32(1e-6), F32(2 * math.pi)
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = np.arange(num_pos_feats, dtype=F32)
    dim_t = (F32(temperature) ** (2 * np.floor(dim_t / 2) / F32(num_pos_feats))).astype(F32)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = np.stack((np.sin(pos_x[..., 0::2]), np.cos(pos_x[..., 1::2])), axis=4).reshape(*pos_x.shape[:3], -1)
    pos_y = np.stack((np.sin(pos_y[..., 0::2]), np.cos(pos_y[..., 1::2])), axis=4).reshape(*pos_y.shape[:3], -1)
    return np.ascontiguousarray(np.concatenate((pos_y, pos_x), axis=3).transpose(0, 3, 1, 2), dtype=F32)


def seq_embedding_sine(max_len: int, d_model: int = 256) -> np.ndarray:
    """SeqEmbeddingSine buffer `te` — vgqa/core/decoder/position_encoding.py:25-41. Returns (max_len,1,d)."""
    position = np.arange(max_len, dtype=F32)[:, None]
    div_term = np.exp(np.arange(0, d_model, 2, dtype=F32) * F32(-math.log(10000.0) / d_model)).astype(F32)
    te = np.zeros((max_len, 1, d_model), F32)
    te[:, 0, 0::2] = np.sin(position * div_term)
    te[:, 0, 1::2] = np.cos(position * div_term)
    return te


def gen_sineembed_for_position(pos_tensor: np.ndarray) -> np.ndarray:
    """vgqa/core/model_utils.py:15-40. pos_tensor (T,B,4) in (cx,cy,w,h) → (T,B,512) ordered (y,x,w,h)."""
    scale = F32(2 * math.pi)
    dim_t = np.arange(128, dtype=F32)
    dim_t = (F32(10000) ** (2 * np.floor(dim_t / 2) / F32(128))).astype(F32)

    def emb(c):
        p = (pos_tensor[:, :, c] * scale)[:, :, None] / dim_t
        return np.stack((np.sin(p[:, :, 0::2]), np.cos(p[:, :, 1::2])), axis=3).reshape(p.shape[0], p.shape[1], -1)

    pos_x, pos_y, pos_w, pos_h = emb(0), emb(1), emb(2), emb(3)
    return np.concatenate((pos_y, pos_x, pos_w, pos_h), axis=2).astype(F32)


# ----------------------------------------------------------------------------------------------
# encoder
# ----------------------------------------------------------------------------------------------
def encoder_layer(sd, p: str, src, pos, mask, nhead: int) -> np.ndarray:
    """TransformerEncoderLayer.forward (post-norm) — vgqa/core/decoder/modal_encoder.py:164-178."""
    qk = src + pos
    a = torch_mha(qk, qk, src, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                  sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], nhead, mask)
    src = layer_norm(src + a, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    h = np.maximum(linear(src, sd[p + "linear1.weight"], sd[p + "linear1.bias"]), 0)
    src = layer_norm(src + linear(h, sd[p + "linear2.weight"], sd[p + "linear2.bias"]),
                     sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    return src


def cross_modal_encoder(sd, vis, vid, pos, text, vis_mask=None, text_mask=None, nhead=8, num_layers=6,
                        prefix="ground_encoder."):
    """CrossModalEncoder.forward + SpatialTemporalEncoder.forward — modal_encoder.py:41-85,115-140.

    vis, vid, pos: (T,256,H,W); text: (L,1,256); vis_mask (T,H,W) bool; text_mask (1,L) bool.
    Returns dict(encoded_feature (S,T,256), encoded_mask (T,S), frames_cls (T,256), videos_cls (256,)).
    """
    T, d, H, W = vis.shape
    P, L = H * W, text.shape[0]
    if vis_mask is None:
        vis_mask = np.zeros((T, H, W), bool)
    if text_mask is None:
        text_mask = np.zeros((1, L), bool)
    vis_mask = vis_mask.copy()
    vis_mask[:, 0, 0] = False  # :46
    f_vis = vis.reshape(T, d, P).transpose(2, 0, 1)
    f_vid = vid.reshape(T, d, P).transpose(2, 0, 1)
    f_pos = pos.reshape(T, d, P).transpose(2, 0, 1)
    vmask = vis_mask.reshape(T, P)
    tmask = np.broadcast_to(text_mask, (T, L))
    ftext = np.broadcast_to(text, (L, T, d))
    x = np.concatenate([f_vis, ftext, f_vid], 0).astype(F32)          # :64
    mask = np.concatenate([vmask, tmask, vmask], 1)                   # :65
    posc = np.concatenate([f_pos, np.zeros_like(ftext), f_pos], 0).astype(F32)  # :66
    for i in range(num_layers):                                       # :125-132 (spatial_layers only)
        x = encoder_layer(sd, f"{prefix}encoder.spatial_layers.{i}.", x, posc, mask, nhead)
    x = layer_norm(x, sd[prefix + "encoder.norm.weight"], sd[prefix + "encoder.norm.bias"])  # :135-136
    frame_src = x.mean(0)                                             # :138
    video_src = frame_src.mean(0)                                     # :139
    return {"encoded_feature": x, "encoded_mask": mask, "frames_cls": frame_src, "videos_cls": video_src,
            "durations": [T], "fea_map_size": (H, W)}


# ----------------------------------------------------------------------------------------------
# classifiers (BERT-style cross-attention blocks)
# ----------------------------------------------------------------------------------------------
def bert_layer_cross(sd, p: str, q: np.ndarray, kv: np.ndarray, nhead: int = 8):
    """BertLayer_Cross.forward — vgqa/core/language/bert_module.py:177-193 (with 34-80, 83-96, 114-141).

    q (B,Lq,256), kv (B,Lk,256) → (layer_output (B,Lq,256), att_map (B,heads,Lq,Lk))."""
    B, Lq, d = q.shape
    Lk = kv.shape[1]
    dh = d // nhead
    mq = linear(q, sd[p + "attention.self.query.weight"], sd[p + "attention.self.query.bias"])
    mk = linear(kv, sd[p + "attention.self.key.weight"], sd[p + "attention.self.key.bias"])
    mv = linear(kv, sd[p + "attention.self.value.weight"], sd[p + "attention.self.value.bias"])
    ql = mq.reshape(B, Lq, nhead, dh).transpose(0, 2, 1, 3)
    kl = mk.reshape(B, Lk, nhead, dh).transpose(0, 2, 1, 3)
    vl = mv.reshape(B, Lk, nhead, dh).transpose(0, 2, 1, 3)
    scores = (ql @ kl.transpose(0, 1, 3, 2)) / F32(math.sqrt(dh))
    probs = softmax(scores, -1)
    ctx = (probs @ vl).transpose(0, 2, 1, 3).reshape(B, Lq, d)
    att_out = bert_layer_norm(
        linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"]) + q,
        sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"])
    inter = gelu_erf(linear(att_out, sd[p + "hidden_intermediate.dense.weight"], sd[p + "hidden_intermediate.dense.bias"]))
    out = bert_layer_norm(
        linear(inter, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"]) + att_out,
        sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"])
    return out, probs


def bert_lm_head(sd, p: str, x: np.ndarray) -> np.ndarray:
    """BertLMPredictionHead — bert_module.py:196-225."""
    h = gelu_erf(linear(x, sd[p + "transform.dense.weight"], sd[p + "transform.dense.bias"]))
    h = bert_layer_norm(h, sd[p + "transform.LayerNorm.weight"], sd[p + "transform.LayerNorm.bias"])
    return linear(h, sd[p + "decoder.weight"]) + sd[p + "bias"]


def temporal_sampling(sd, p: str, x: np.ndarray, query: np.ndarray) -> np.ndarray:
    """TemporalSampling.forward — vgqa/core/decoder/classifier.py:32-37.

    x (T,256,H,W) → pooled (1,T,256) as *queries*; `query` = text (1,L,256) as key/value. Returns (T,)."""
    h = x.mean(axis=(2, 3))[None]  # adaptive_avg_pool2d → (1,T,256)
    for i in range(2):
        h, _ = bert_layer_cross(sd, f"{p}layer_ca.{i}.", h, query)
    return bert_lm_head(sd, p + "head.", h).reshape(-1)


def spatial_activation(sd, p: str, inp: np.ndarray, init_q: np.ndarray):
    """SpatialActivation.forward — classifier.py:64-81.

    inp (K,256,H,W), init_q (1,1,256) → logits (1,vocab), att_map (K,P)."""
    K = inp.shape[0]
    x = inp.transpose(0, 2, 3, 1).reshape(K, -1, 256)
    query = np.repeat(init_q, K, axis=0)
    att = None
    for i in range(2):
        query, att = bert_layer_cross(sd, f"{p}layer_ca.{i}.", query, x)
    att_map = sigmoid(att.sum(1)[:, 0, :])                       # :75  (K,P)
    amin = att_map.min(1, keepdims=True)
    amax = att_map.max(1, keepdims=True)
    att_map = (att_map - amin) / (amax - amin + F32(1e-6))        # :76-78
    logits = bert_lm_head(sd, p + "head.", query).mean(0)         # :80  (1,vocab)
    return logits.astype(F32), att_map.astype(F32)


# ----------------------------------------------------------------------------------------------
# decoder
# ----------------------------------------------------------------------------------------------
def custom_mha(q, k, v, out_w, out_b, nhead: int) -> np.ndarray:
    """Projection-free MultiheadAttention(512, 8, vdim=256) — vgqa/core/decoder/attention.py:116-260.

    q (1,bs,512), k (n,bs,512), v (n,bs,256); q *= (512/8)**-0.5; explicit max-subtracted softmax."""
    tgt_len, bsz, e = q.shape
    dh = e // nhead
    vdh = v.shape[2] // nhead
    qh = (q * F32(dh ** -0.5)).reshape(tgt_len, bsz * nhead, dh).transpose(1, 0, 2)
    kh = k.reshape(-1, bsz * nhead, dh).transpose(1, 0, 2)
    vh = v.reshape(-1, bsz * nhead, vdh).transpose(1, 0, 2)
    s = qh @ kh.transpose(0, 2, 1)
    p = softmax(s - s.max(-1, keepdims=True), -1)
    o = (p @ vh).transpose(1, 0, 2).reshape(tgt_len, bsz, v.shape[2])
    return linear(o, out_w, out_b)


def time_decoder(sd, tgt, query_time, query_mask, mem, mem_pos, mem_mask, nhead=8, num_layers=6,
                 prefix="ground_decoder.time_decoder."):
    """TimeDecoder / TimeDecoderLayer — vgqa/core/decoder/query_decoder.py:379-423,456-486.

    tgt (T,1,256); mem (M,T,256); returns stacked normed intermediates (6,1,T,256)."""
    inter = []
    for i in range(num_layers):
        p = f"{prefix}layers.{i}."
        qk = tgt + query_time
        t2 = torch_mha(qk, qk, tgt, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                       sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], nhead, query_mask)
        tgt = layer_norm(tgt + t2, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        t2 = torch_mha(tgt.transpose(1, 0, 2), mem + mem_pos, mem,
                       sd[p + "cross_attn_image.in_proj_weight"], sd[p + "cross_attn_image.in_proj_bias"],
                       sd[p + "cross_attn_image.out_proj.weight"], sd[p + "cross_attn_image.out_proj.bias"],
                       nhead, mem_mask)
        tgt = layer_norm(tgt + t2.transpose(1, 0, 2), sd[p + "norm3.weight"], sd[p + "norm3.bias"])
        h = np.maximum(linear(tgt, sd[p + "linear1.weight"], sd[p + "linear1.bias"]), 0)
        tgt = layer_norm(tgt + linear(h, sd[p + "linear2.weight"], sd[p + "linear2.bias"]),
                         sd[p + "norm4.weight"], sd[p + "norm4.bias"])
        inter.append(layer_norm(tgt, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"]))  # :412 shared norm
    return np.stack(inter).transpose(0, 2, 1, 3)  # (6,1,T,256)


def pos_decoder(sd, tgt, boxes, query_time, mem, mem_pos, nhead=8, num_layers=6,
                prefix="ground_decoder.decoder.", bbox_prefix="bbox_embed"):
    """PosDecoder / PosDecoderLayer — query_decoder.py:151-205,266-375 (FROM_SCRATCH=True branch).

    tgt (T,1,256); boxes (T,1,4) sigmoid-ed anchors; mem (M,T,256) = [vis‖text]; returns (6,1,T,4)."""
    d = 256
    anchors = []
    for lid in range(num_layers):
        p = f"{prefix}layers.{lid}."
        sine = gen_sineembed_for_position(boxes)                                      # :169
        query_pos = mlp(sine, sd, prefix + "ref_point_head", 2)                       # :170
        pos_tr = F32(1.0) if lid == 0 else mlp(tgt, sd, prefix + "query_scale", 2)     # :173-176
        sine = sine[..., :d] * pos_tr                                                 # :179

        def L(name, x):
            return linear(x, sd[p + name + ".weight"], sd[p + name + ".bias"])

        # ---- temporal self-attention (:282-296) ----
        q = L("sa_qcontent_proj", tgt) + L("sa_qtime_proj", query_time) + L("sa_qpos_proj", query_pos)
        k = L("sa_kcontent_proj", tgt) + L("sa_ktime_proj", query_time) + L("sa_kpos_proj", query_pos)
        v = L("sa_v_proj", tgt)
        t2 = torch_mha(q, k, v, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                       sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], nhead)
        x = layer_norm(tgt + t2, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        # ---- conditional cross-attention (:301-369) ----
        t, b, c = x.shape
        n_tok, bs, f = mem.shape
        q_content = L("ca_qcontent_proj", x)
        k_content = L("ca_kcontent_proj", mem)
        vv = L("ca_v_proj", mem)
        k_pos = L("ca_kpos_proj", mem_pos)
        if lid == 0:
            qq = q_content + L("ca_qpos_proj", query_pos)
            kk = k_content + k_pos
        else:
            qq, kk = q_content, k_content
        qq = qq.reshape(t, b, nhead, c // nhead)
        sp = L("ca_qpos_sine_proj", sine).reshape(t, b, nhead, c // nhead)
        qq = np.concatenate([qq, sp], 3).reshape(t, b, 2 * c)
        kk = kk.reshape(n_tok, bs, nhead, f // nhead)
        kp = k_pos.reshape(n_tok, bs, nhead, f // nhead)
        kk = np.concatenate([kk, kp], 3).reshape(n_tok, bs, 2 * f)
        q_cross = qq[:, 0, :][None]                                                   # (1, T, 512) :340-345
        t2 = custom_mha(q_cross, kk, vv, sd[p + "cross_attn.out_proj.weight"], sd[p + "cross_attn.out_proj.bias"], nhead)
        t2 = t2.reshape(b, t, f).transpose(1, 0, 2)                                   # :361-366
        x = layer_norm(x + t2, sd[p + "norm3.weight"], sd[p + "norm3.bias"])
        h = np.maximum(L("linear1", x), 0)
        x = layer_norm(x + L("linear2", h), sd[p + "norm4.weight"], sd[p + "norm4.bias"])
        tgt = x
        boxes = sigmoid(mlp(tgt, sd, bbox_prefix, 3))                                 # :188-192
        anchors.append(boxes)
    return np.stack(anchors).transpose(0, 2, 1, 3)  # (6,1,T,4)  :205


def query_decoder(sd, enc, vis_pos, itq, isq, nhead=8, num_layers=6, prefix="ground_decoder."):
    """QueryDecoder.forward — query_decoder.py:76-126. Note the keyword order quirk: called with
    isq=spatial query, itq=temporal query (grounding_net.py:139-141)."""
    feat = enc["encoded_feature"]
    H, W = enc["fea_map_size"]
    P = H * W
    T = feat.shape[1]
    d = feat.shape[2]
    encoded_pos = vis_pos.reshape(T, d, P).transpose(2, 0, 1)
    zeros_txt = np.zeros_like(feat[P:-P])
    pos_s = np.concatenate([encoded_pos, zeros_txt], 0)     # :82
    pos_t = np.concatenate([zeros_txt, encoded_pos], 0)     # :83
    fc = enc["frames_cls"]
    # pos_fc: BertLN(256) → Linear(256,4) → ReLU → BertLN(4)  (:53-59, :92-94)
    h = bert_layer_norm(fc, sd[prefix + "pos_fc.0.weight"], sd[prefix + "pos_fc.0.bias"])
    h = np.maximum(linear(h, sd[prefix + "pos_fc.2.weight"], sd[prefix + "pos_fc.2.bias"]), 0)
    h = bert_layer_norm(h, sd[prefix + "pos_fc.4.weight"], sd[prefix + "pos_fc.4.bias"])
    pos_query = sigmoid(h)[:, None, :]                      # (T,1,4)
    # time_fc(videos_cls) is computed by the reference but never consumed (:456-486) — omitted.
    te = sd[prefix + "time_embed.te"][:T]                   # (T,1,256)  :98
    query_mask = np.zeros((1, T), bool)
    mem_mask = enc["encoded_mask"][:, :-P]                  # :100
    tgt_t = np.broadcast_to(itq[None, None, :], (T, 1, d)).astype(F32)
    out_time = time_decoder(sd, tgt_t, te, query_mask, feat[P:], pos_t, mem_mask, nhead, num_layers,
                            prefix + "time_decoder.")
    tgt_s = np.broadcast_to(isq[None, None, :], (T, 1, d)).astype(F32)
    out_pos = pos_decoder(sd, tgt_s, pos_query, te, feat[:-P], pos_s, nhead, num_layers,
                          prefix + "decoder.", "bbox_embed")
    return out_pos, out_time


# ----------------------------------------------------------------------------------------------
# the hot path: VSTGNet.forward after the feature extractors
# ----------------------------------------------------------------------------------------------
def _choose(att_bool: np.ndarray, fallback: np.ndarray) -> List[int]:
    idx = np.nonzero(att_bool)[0].tolist()
    return idx or np.nonzero(fallback)[0].tolist()


def hot_path_forward(sd, vis, vid, pos, text, vis_mask=None, text_mask=None, iteration_rate: int = -1,
                     theta: float = 0.45, nhead=8, enc_layers=6, dec_layers=6, return_debug=False):
    """VSTGNet.forward lines 114-202 — vgqa/core/grounding_net.py — on post-`input_proj` features.

    vis = input_proj(resnet feats) (T,256,H,W); vid = input_proj2(swin feats) (T,256,H,W);
    pos = PositionEmbeddingSine (T,256,H,W); text = resizer(RoBERTa) tokens (L,1,256)."""
    T, d, H, W = vis.shape
    P = H * W
    enc = cross_modal_encoder(sd, vis, vid, pos, text, vis_mask, text_mask, nhead, enc_layers)
    feat = enc["encoded_feature"]
    f_vid = feat[-P:].transpose(1, 2, 0).reshape(T, d, H, W)       # :117
    f_vis = feat[:P].transpose(1, 2, 0).reshape(T, d, H, W)        # :118
    f_text_cls = feat[P:-P].mean(1)[None]                          # :119 (1,L,256)
    logits_f_m = temporal_sampling(sd, "t_temporal_clas.", f_vid, f_text_cls)   # :122
    logits_f_a = temporal_sampling(sd, "s_temporal_clas.", f_vis, f_text_cls)   # :123
    att = (sigmoid(logits_f_m) + sigmoid(logits_f_a)) / F32(2)                   # :125
    choose = _choose(att > F32(theta), att > 0)                                   # :126-128
    dbg = {"choose_pass1": list(choose)}

    def seed(choose):
        lr_m, am_t = spatial_activation(sd, "t_spatial_clas.", f_vid[choose], f_text_cls[:, :1])
        lr_a, am_s = spatial_activation(sd, "s_spatial_clas.", f_vis[choose], f_text_cls[:, :1])
        itq = (feat[-P:].transpose(1, 0, 2)[choose] * am_t[:, :, None]).mean((0, 1))   # :135
        isq = (feat[:P].transpose(1, 0, 2)[choose] * am_s[:, :, None]).mean((0, 1))    # :136
        return lr_m, lr_a, itq.astype(F32), isq.astype(F32)

    logits_r_m, logits_r_a, itq, isq = seed(choose)
    out_pos, out_time = query_decoder(sd, enc, pos, itq, isq, nhead, dec_layers)        # :139-141
    if iteration_rate < 0:                                                               # :143-163
        act = sigmoid(mlp(out_time, sd, "action_embed", 2)[-1].reshape(-1))
        dbg["actioness_pass1"] = act
        choose = _choose(act > F32(0.5), att > 0)
        dbg["choose_pass2"] = list(choose)
        logits_r_m, logits_r_a, itq, isq = seed(choose)
        out_pos, out_time = query_decoder(sd, enc, pos, itq, isq, nhead, dec_layers)
    coord = out_pos.reshape(out_pos.shape[0], -1, 4)               # flatten(1,2) :168
    sted = mlp(out_time, sd, "temp_embed", 2)                      # :178
    actn = mlp(out_time, sd, "action_embed", 2)                    # :179
    out = {
        "pred_boxes": coord[-1], "logits_f_m": logits_f_m, "logits_f_a": logits_f_a,
        "logits_r_a": logits_r_a, "logits_r_m": logits_r_m,
        "pred_sted": sted[-1], "pred_actioness": actn[-1],
        "aux_outputs": [{"pred_sted": a, "pred_boxes": b, "pred_actioness": c}
                        for a, b, c in zip(sted[:-1], coord[:-1], actn[:-1])],
        "att_sequences": att[None],
        "choose_index": list(choose),
    }
    if return_debug:
        dbg["encoded_feature"] = feat
        dbg["frames_cls"] = enc["frames_cls"]
        out["debug"] = dbg
    return out


# ----------------------------------------------------------------------------------------------
# decode: PostProcess + tube/segment assembly
# ----------------------------------------------------------------------------------------------
def box_cxcywh_to_xyxy(x: np.ndarray) -> np.ndarray:
    """vgqa/utils/box_ops.py:44-51."""
    cx, cy, w, h = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    return np.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1).astype(F32)


def postprocess(pred_boxes, pred_sted, att_sequences, target_sizes, frames_id, durations):
    """PostProcess.forward — vgqa/core/postprocessor.py:14-50.

    pred_boxes (T,4) cxcywh; pred_sted (b,T,2); target_sizes (T,2) = (h,w). Returns
    (boxes_xyxy_px (T,4), att (b,T), [[start_fid, end_fid+1]]) and the argmax (start_idx,end_idx)."""
    assert len(pred_boxes) == len(target_sizes)
    boxes = box_cxcywh_to_xyxy(pred_boxes)
    img_h, img_w = target_sizes[:, 0].astype(F32), target_sizes[:, 1].astype(F32)
    scale = np.stack([img_w, img_h, img_w, img_h], 1)
    boxes = np.maximum(boxes * scale, 0).astype(F32)
    b, t, _ = pred_sted.shape
    inf = F32(-1e32)
    steds, idxs = [], []
    for i_b, duration in enumerate(durations):
        m = np.tril(np.full((t, t), inf, F32), 0)
        m[duration:, :] = inf
        m[:, duration:] = inf
        pm = m + log_softmax(pred_sted[i_b, :, 0], 0)[:, None] + log_softmax(pred_sted[i_b, :, 1], 0)[None, :]
        k = int(np.argmax(pm.reshape(-1)))
        s, e = k // t, k % t
        idxs.append((s, e))
        steds.append([frames_id[i_b][s], frames_id[i_b][e] + 1])
    return boxes, att_sequences, steds, idxs


def linear_interp(bbox_dict: Dict[int, List[List[float]]]):
    """vgqa/training/evaluator.py:10-36."""
    fids = sorted(bbox_dict.keys())
    if len(fids) < 2:
        return bbox_dict
    for i in range(len(fids) - 1):
        l, r = fids[i], fids[i + 1]
        if r - l > 1:
            n = r - l
            bl, br = bbox_dict[l][0], bbox_dict[r][0]
            dl = [(br[j] - bl[j]) / n for j in range(4)]
            for step in range(1, n):
                bbox_dict[l + step] = [[bl[j] + step * dl[j] for j in range(4)]]
    fids = sorted(bbox_dict.keys())
    assert max(fids) - min(fids) + 1 == len(fids)
    return {f: bbox_dict[f] for f in fids}


def linear_interp_conf(conf_dict):
    """vgqa/training/evaluator.py:39-54 (nearest-hold: left if step <= interval//2 else right)."""
    fids = sorted(conf_dict.keys())
    if len(fids) < 2:
        return conf_dict
    for i in range(len(fids) - 1):
        l, r = fids[i], fids[i + 1]
        if r - l > 1:
            n = r - l
            for step in range(1, n):
                conf_dict[l + step] = conf_dict[l] if step <= (n // 2) else conf_dict[r]
    fids = sorted(conf_dict.keys())
    assert max(fids) - min(fids) + 1 == len(fids)
    return {f: conf_dict[f] for f in fids}


def single_forward_dicts(out, ori_size: Tuple[int, int], frame_ids: Sequence[int], vid_key=0):
    """single_forward — vgqa/training/evaluator.py:56-92, for one clip: returns (bbox_pred, att_pred, sted)."""
    T = out["pred_boxes"].shape[0]
    sizes = np.tile(np.asarray(ori_size, F32)[None], (T, 1))
    boxes, att, steds, _ = postprocess(out["pred_boxes"], out["pred_sted"], out["att_sequences"], sizes,
                                       [list(frame_ids)], [T])
    bbox_pred = {frame_ids[i]: [boxes[i].tolist()] for i in range(T)}
    att_pred = {frame_ids[i]: [float(att[0][i])] for i in range(T)}
    return bbox_pred, att_pred, steds[0]


def merge_predict(pass1, pass2, fps: float):
    """predict() merge — vgqa/inference/grounding.py:214-244. pass_k = (bbox_pred, att_pred, sted)."""
    bbox = dict(pass1[0]); bbox.update(pass2[0])
    bbox_full = linear_interp(bbox)
    att = dict(pass1[1]); att.update(pass2[1])
    att_full = linear_interp_conf(att)
    sted = [min(pass1[2][0], pass2[2][0]), max(pass1[2][1], pass2[2][1])]
    temporal = {"start": float(sted[0]) / max(fps, 1e-6), "end": float(sted[1]) / max(fps, 1e-6), "score": 1.0}
    tube = []
    for fid in sorted(bbox_full.keys()):
        b = bbox_full[fid][0]
        c = att_full.get(fid, 1.0)
        tube.append({"frame": int(fid), "bbox": [float(b[0]), float(b[1]), float(b[2]), float(b[3])],
                     "score": float(c[0] if isinstance(c, list) else c)})
    return {"temporal": temporal, "tube": tube}


def input_proj(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """1x1 Conv2d — grounding_net.py:62,71,101,105. x (T,C,H,W), w (256,C,1,1) → (T,256,H,W)."""
    T, C, H, W = x.shape
    y = x.transpose(0, 2, 3, 1).reshape(-1, C) @ w.reshape(w.shape[0], C).T + b
    return y.reshape(T, H, W, -1).transpose(0, 3, 1, 2).astype(F32)


def feature_resizer(x: np.ndarray, sd: Dict[str, np.ndarray], prefix: str = "text_encoder.resizer") -> np.ndarray:
    """FeatureResizer — vgqa/core/language/bert.py:77-96: fc → nn.LayerNorm(eps=1e-12) → dropout (identity in eval)."""
    y = linear(x, sd[prefix + ".fc.weight"], sd[prefix + ".fc.bias"])
    return layer_norm(y, sd[prefix + ".layer_norm.weight"], sd[prefix + ".layer_norm.bias"], 1e-12)


def roberta_encoder(sd, input_ids: np.ndarray, attention_pad: Optional[np.ndarray] = None, prefix: str = "text_encoder.body.",
                    pad_id: int = 1, eps: float = 1e-5) -> np.ndarray:
    """RoBERTa-base encoder → `last_hidden_state` (bert.py:66,69: `self.body(**tokenized).last_hidden_state`).

    The arithmetic lives in a third-party dependency that is NOT vendored in the reference: `transformers` (pinned 4.37.2 in
    the reference's requirements.txt:2; 5.5.0 in this image) `RobertaModel` = `RobertaEmbeddings` + 12 x `RobertaLayer`.
    Restated from its published algorithm: position ids = cumsum(ids != pad) * (ids != pad) + pad; embeddings = word +
    token_type[0] + position → LayerNorm(eps 1e-5 in roberta-base's config) ; per layer: q/k/v Linear, heads of 64,
    softmax(q k^T / 8 + key padding mask) v, dense + LayerNorm(x + ·), dense 3072 erf-GELU, dense + LayerNorm(x + ·).
    Pinned by tests/golden/make_golden_text.py against transformers' own RobertaModel run in this container.

    input_ids (B, L) int; attention_pad (B, L) bool, True = padded (the reference's `attention_mask.ne(1)`, bert.py:70)."""
    ids = np.asarray(input_ids)
    B, L = ids.shape
    keep = ids != pad_id
    pos_ids = np.cumsum(keep, axis=1) * keep + pad_id
    e = prefix + "embeddings."
    x = sd[e + "word_embeddings.weight"][ids] + sd[e + "position_embeddings.weight"][pos_ids] + sd[e + "token_type_embeddings.weight"][0]
    x = layer_norm(x, sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], eps)
    d = x.shape[-1]
    nh, dh = d // 64, 64
    i = 0
    while f"{prefix}encoder.layer.{i}.attention.self.query.weight" in sd:
        p = f"{prefix}encoder.layer.{i}."
        lin = lambda t, n: linear(t, sd[p + n + ".weight"], sd[p + n + ".bias"])
        q = lin(x, "attention.self.query").reshape(B, L, nh, dh).transpose(0, 2, 1, 3)
        k = lin(x, "attention.self.key").reshape(B, L, nh, dh).transpose(0, 2, 1, 3)
        v = lin(x, "attention.self.value").reshape(B, L, nh, dh).transpose(0, 2, 1, 3)
        sc = (q @ k.transpose(0, 1, 3, 2)) / F32(math.sqrt(dh))
        if attention_pad is not None:
            sc = np.where(np.asarray(attention_pad, bool)[:, None, None, :], -np.inf, sc)
        ctx = (softmax(sc, -1) @ v).transpose(0, 2, 1, 3).reshape(B, L, d)
        a = layer_norm(lin(ctx, "attention.output.dense") + x, sd[p + "attention.output.LayerNorm.weight"],
                       sd[p + "attention.output.LayerNorm.bias"], eps)
        h = gelu_erf(lin(a, "intermediate.dense"))
        x = layer_norm(lin(h, "output.dense") + a, sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], eps)
        i += 1
    return x.astype(F32)


def front_end(sd, vis_raw: np.ndarray, vid_raw: np.ndarray, text_raw: np.ndarray):
    """The step right before the hot path: grounding_net.py:101 `input_proj(vis_res_features)`,
    :105 `input_proj2(vid_features_all['3'])`, bert.py:70,73 `resizer(last_hidden_state.transpose(0, 1))`.

    vis_raw (T,Cv,H,W), vid_raw (T,Cd,H,W), text_raw (L,Ct) → vis (T,256,H,W), vid (T,256,H,W), text (L,1,256)."""
    vis = input_proj(vis_raw, sd["input_proj.weight"], sd["input_proj.bias"])
    vid = input_proj(vid_raw, sd["input_proj2.weight"], sd["input_proj2.bias"])
    text = feature_resizer(text_raw, sd)[:, None, :]
    return vis, vid, text


# ----------------------------------------------------------------------------------------------
# widened row: the last stage of the Video-Swin-T extractor
# ----------------------------------------------------------------------------------------------
def swin_relative_position_index(window: Tuple[int, int, int]) -> np.ndarray:
    """WindowAttention3D.__init__ — vgqa/core/vision/video_swin_transformer.py:97-112: (N, N) indices into the bias table."""
    wd, wh, ww = window
    coords = np.stack(np.meshgrid(np.arange(wd), np.arange(wh), np.arange(ww), indexing="ij")).reshape(3, -1)
    rel = (coords[:, :, None] - coords[:, None, :]).transpose(1, 2, 0).copy()
    rel[:, :, 0] += wd - 1
    rel[:, :, 1] += wh - 1
    rel[:, :, 2] += ww - 1
    rel[:, :, 0] *= (2 * wh - 1) * (2 * ww - 1)
    rel[:, :, 1] *= (2 * ww - 1)
    return rel.sum(-1)


def _swin_windows(x: np.ndarray, win) -> np.ndarray:
    """window_partition (video_swin_transformer.py:36-42): (B, D, H, W, C) → (B*nW, wd*wh*ww, C)."""
    B, D, H, W, C = x.shape
    wd, wh, ww = win
    x = x.reshape(B, D // wd, wd, H // wh, wh, W // ww, ww, C).transpose(0, 1, 3, 5, 2, 4, 6, 7)
    return x.reshape(-1, wd * wh * ww, C)


def _swin_unwindows(w: np.ndarray, win, B, D, H, W) -> np.ndarray:
    """window_reverse (:45-50)."""
    wd, wh, ww = win
    x = w.reshape(B, D // wd, H // wh, W // ww, wd, wh, ww, -1).transpose(0, 1, 4, 2, 5, 3, 6, 7)
    return x.reshape(B, D, H, W, -1)


def swin_shift_mask(D, H, W, win, shift) -> np.ndarray:
    """compute_mask (:311-325): (nW, N, N) of 0 / -100 for the rolled map."""
    img = np.zeros((D, H, W), np.int64)
    cnt = 0
    for d in (slice(-win[0]), slice(-win[0], -shift[0]), slice(-shift[0], None)):
        for h in (slice(-win[1]), slice(-win[1], -shift[1]), slice(-shift[1], None)):
            for w in (slice(-win[2]), slice(-win[2], -shift[2]), slice(-shift[2], None)):
                img[d, h, w] = cnt
                cnt += 1
    mw = _swin_windows(img[None, ..., None].astype(F32), win)[..., 0]
    return np.where(mw[:, None, :] != mw[:, :, None], F32(-100.0), F32(0.0))


def swin_stage(sd, x: np.ndarray, prefix: str = "vid.layers.3.", heads: int = 24, window=(8, 7, 7), depth: int = 2) -> np.ndarray:
    """BasicLayer.forward of one Video-Swin stage (no downsample) — video_swin_transformer.py:377-398 — on a channels-last map
    x (B, D, H, W, C):
    per block (:200-275)  x += proj(W-MSA(LN1(x)));  x += fc2(gelu(fc1(LN2(x)))), window attention with the relative position bias
    (:143-165), odd blocks on the cyclically rolled map with the -100 mask of compute_mask (:311-325).  Sides that are not multiples
    of the window: LN1(x) is zero-padded at the END of every axis up to the next multiple (:205-211; the padded tokens take part in
    the attention as keys with q = k = v = the biases), the mask is computed on the padded sides (:383-386), the result is cropped
    (:233-234)."""
    B, D, H, W, C = x.shape
    size = (D, H, W)
    win = tuple(min(window[i], size[i]) for i in range(3))           # get_window_size (:53-66)
    shift = tuple(0 if size[i] <= window[i] else window[i] // 2 for i in range(3))
    Dp, Hp, Wp = (int(math.ceil(size[i] / win[i])) * win[i] for i in range(3))
    N = win[0] * win[1] * win[2]
    dh = C // heads
    idx = swin_relative_position_index(window)
    # tokens of a clamped window: the first win[i] coordinates of every axis of the configured window (index[:N, :N] in the
    # reference is only the same thing when the window is not clamped in H or W)
    assert win[1:] == tuple(window[1:]) or win == tuple(window), "oracle: spatially clamped windows are not restated"
    idx = idx[:N, :N]
    mask = swin_shift_mask(Dp, Hp, Wp, win, shift) if any(shift) else None
    x = x.astype(F32)
    for i in range(depth):
        p = f"{prefix}blocks.{i}."
        sh = shift if (i % 2 == 1 and any(shift)) else None
        h = layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
        h = np.pad(h, ((0, 0), (0, Dp - D), (0, Hp - H), (0, Wp - W), (0, 0)))
        if sh:
            h = np.roll(h, (-sh[0], -sh[1], -sh[2]), axis=(1, 2, 3))
        winx = _swin_windows(h, win)
        qkv = linear(winx, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(-1, N, 3, heads, dh).transpose(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * F32(dh ** -0.5), qkv[1], qkv[2]
        attn = q @ k.transpose(0, 1, 3, 2)
        bias = sd[p + "attn.relative_position_bias_table"][idx.reshape(-1)].reshape(N, N, heads).transpose(2, 0, 1)
        attn = attn + bias[None]
        if sh:
            nW = mask.shape[0]
            attn = (attn.reshape(B, nW, heads, N, N) + mask[None, :, None]).reshape(-1, heads, N, N)
        attn = softmax(attn, -1)
        o = (attn @ v).transpose(0, 2, 1, 3).reshape(-1, N, C)
        o = _swin_unwindows(linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]), win, B, Dp, Hp, Wp)
        if sh:
            o = np.roll(o, sh, axis=(1, 2, 3))
        x = x + o[:, :D, :H, :W]
        h = layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        h = gelu_erf(linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        x = x + linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x.astype(F32)


def swin_patch_merging(sd, x: np.ndarray, prefix: str) -> np.ndarray:
    """PatchMerging.forward (:291-308) for even H, W: 2x2 neighbours concatenated (x0, x1, x2, x3 order), LayerNorm(4C), Linear(4C → 2C)."""
    x0, x1, x2, x3 = x[:, :, 0::2, 0::2], x[:, :, 1::2, 0::2], x[:, :, 0::2, 1::2], x[:, :, 1::2, 1::2]
    h = np.concatenate([x0, x1, x2, x3], -1)
    h = layer_norm(h, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])
    return linear(h, sd[prefix + "reduction.weight"])


def video_swin_backbone(sd, frames: np.ndarray, clips: int, prefix: str = "vid.", depths=(2, 2, 6, 2), heads=(3, 6, 12, 24),
                        window=(8, 7, 7)):
    """VideoSwinTransformerBackbone.forward (:666-685) on frames (clips*T, 3, R, R), R a multiple of 4: PatchEmbed3D with patch
    (1,4,4) + LayerNorm (:426-443), four stages with PatchMerging between them.  Returns the four stage outputs as channels-last
    maps [(clips, T, H_s, W_s, C_s)] ('0'..'3' of the reference, which are (clips*T, C, H, W))."""
    n, c, R, _ = frames.shape
    T = n // clips
    w = sd[prefix + "patch_embed.proj.weight"]                      # (96, 3, 1, 4, 4)
    E = w.shape[0]
    x = frames.reshape(clips, T, c, R // 4, 4, R // 4, 4).transpose(0, 1, 3, 5, 2, 4, 6).reshape(clips, T, R // 4, R // 4, c * 16)
    x = linear(x.astype(F32), w.reshape(E, -1), sd[prefix + "patch_embed.proj.bias"])
    x = layer_norm(x, sd[prefix + "patch_embed.norm.weight"], sd[prefix + "patch_embed.norm.bias"])
    outs = []
    for s in range(len(depths)):
        x = swin_stage(sd, x, f"{prefix}layers.{s}.", heads[s], window, depths[s])
        outs.append(x)
        if s + 1 < len(depths):
            x = swin_patch_merging(sd, x, f"{prefix}downsamples.{s}.")
    return outs


# ----------------------------------------------------------------------------------------------
# ResNet101 extractor.  The network itself is torchvision's (third-party, not vendored in the reference: torchvision 0.26.0 in
# this image; `torchvision.models.resnet.ResNet._forward_impl` / `Bottleneck.forward`, v1.5 = the stride sits on the 3x3
# convolution); the reference's part is the construction (backbone.py:104-113: resnet101, norm_layer=FrozenBatchNorm2d, no
# dilation, IntermediateLayerGetter → layer4) and FrozenBatchNorm2d (:13-57).  Pinned by tests/golden/resnet101_*.npz
# (generated from exactly that construction by tests/golden/make_golden_resnet.py).
# ----------------------------------------------------------------------------------------------
def conv2d(x: np.ndarray, w: np.ndarray, stride: int = 1, pad: int = 0) -> np.ndarray:
    """F.conv2d without bias: x (n, C, H, W), w (O, C, k, k)."""
    k = w.shape[2]
    if pad:
        x = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    win = np.lib.stride_tricks.sliding_window_view(x, (k, k), axis=(2, 3))[:, :, ::stride, ::stride]   # (n, C, Ho, Wo, k, k)
    return np.einsum("nchwyx,ocyx->nohw", win, w, optimize=True).astype(F32)


def frozen_bn(x: np.ndarray, sd, p: str) -> np.ndarray:
    """FrozenBatchNorm2d.forward (backbone.py:47-57): eps = 1e-5, scale = weight * rsqrt(var + eps), bias = bias - mean * scale."""
    scale = sd[p + ".weight"] / np.sqrt(sd[p + ".running_var"] + 1e-5)
    bias = sd[p + ".bias"] - sd[p + ".running_mean"] * scale
    return (x * scale[None, :, None, None] + bias[None, :, None, None]).astype(F32)


def max_pool_3x3_s2(x: np.ndarray) -> np.ndarray:
    """nn.MaxPool2d(kernel_size=3, stride=2, padding=1)."""
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)), constant_values=-np.inf)
    win = np.lib.stride_tricks.sliding_window_view(xp, (3, 3), axis=(2, 3))[:, :, ::2, ::2]
    return win.max(axis=(-1, -2)).astype(F32)


def resnet101_backbone(sd, frames: np.ndarray, prefix: str = "vis_encoder.0.body.", blocks=(3, 4, 23, 3)):
    """`Backbone.body` (backbone.py:80-84,104-113) on frames (n, 3, R, R): stem (7x7 / 2 conv, FrozenBN, ReLU, 3x3 / 2 max-pool),
    then the Bottleneck layers.  Returns the outputs of layer1..4 as channels-last maps [(n, H_l, W_l, C_l)]; the reference keeps
    layer4 (n, 2048, R/32, R/32)."""
    relu = lambda v: np.maximum(v, 0.0)
    x = relu(frozen_bn(conv2d(frames.astype(F32), sd[prefix + "conv1.weight"], 2, 3), sd, prefix + "bn1"))
    x = max_pool_3x3_s2(x)
    outs = []
    for l, nb in enumerate(blocks):
        for b in range(nb):
            q = f"{prefix}layer{l + 1}.{b}."
            stride = 2 if (b == 0 and l > 0) else 1
            o = relu(frozen_bn(conv2d(x, sd[q + "conv1.weight"]), sd, q + "bn1"))
            o = relu(frozen_bn(conv2d(o, sd[q + "conv2.weight"], stride, 1), sd, q + "bn2"))
            o = frozen_bn(conv2d(o, sd[q + "conv3.weight"]), sd, q + "bn3")
            idn = x
            if b == 0:
                idn = frozen_bn(conv2d(x, sd[q + "downsample.0.weight"], stride, 0), sd, q + "downsample.1")
            x = relu(o + idn)
        outs.append(np.ascontiguousarray(x.transpose(0, 2, 3, 1)))
    return outs


# ----------------------------------------------------------------------------------------------
# deterministic synthetic weights / inputs: they live in the product package (pure numpy, vgqa_b200/synth.py) so that the
# GPU arm of bench.py imports nothing from oracle/; re-exported here for the tests and fixture makers
# ----------------------------------------------------------------------------------------------
from vgqa_b200.synth import (CALIB_PREFIX, apply_calibration, hot_path_param_shapes, synth_event_inputs, synth_inputs,  # noqa: E402,F401
                             synth_masks, synth_raw_inputs, synth_resnet101, synth_state_dict, synth_swin_backbone, synth_swin_stage, synth_text_ids)
