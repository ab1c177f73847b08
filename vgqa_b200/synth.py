"""Deterministic synthetic weights and inputs (numpy PCG64 — identical on every machine, no torch RNG).

Shared by the golden-fixture makers, the tests, bench.py and smoke(): the 36 M hot-path weights are regenerated from a
seed instead of being stored.  Pure numpy: it depends neither on the CPU checker used by the tests nor on the CUDA
library, so the GPU arm of bench.py can import it alone.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np

F32 = np.float32


def seq_embedding_sine(max_len: int, d_model: int = 256) -> np.ndarray:
    """The `time_embed.te` buffer (SeqEmbeddingSine, vgqa/core/decoder/position_encoding.py:25-41): (max_len,1,d)."""
    position = np.arange(max_len, dtype=F32)[:, None]
    div_term = np.exp(np.arange(0, d_model, 2, dtype=F32) * F32(-math.log(10000.0) / d_model)).astype(F32)
    te = np.zeros((max_len, 1, d_model), F32)
    te[:, 0, 0::2] = np.sin(position * div_term)
    te[:, 0, 1::2] = np.cos(position * div_term)
    return te


def sine_position_table(T: int, H: int, W: int) -> np.ndarray:
    """PositionEmbeddingSine(128, normalize=True) of an all-False (T,H,W) mask (vision/position_encoding.py:50-91):
    (T,256,H,W) fp32, identical for every frame."""
    y = np.broadcast_to(np.arange(1, H + 1, dtype=F32)[:, None], (H, W))
    x = np.broadcast_to(np.arange(1, W + 1, dtype=F32)[None, :], (H, W))
    eps, scale = F32(1e-6), F32(2 * math.pi)
    y = y / (F32(H) + eps) * scale
    x = x / (F32(W) + eps) * scale
    dim_t = np.arange(128, dtype=F32)
    dim_t = (F32(10000.0) ** (2 * np.floor(dim_t / 2) / F32(128))).astype(F32)

    def emb(e):
        p = e[:, :, None] / dim_t
        return np.stack((np.sin(p[..., 0::2]), np.cos(p[..., 1::2])), axis=3).reshape(H, W, -1)

    pos = np.concatenate((emb(y), emb(x)), axis=2).transpose(2, 0, 1)
    return np.ascontiguousarray(np.broadcast_to(pos[None], (T, 256, H, W)), dtype=F32)

def hot_path_param_shapes(enc_layers=6, dec_layers=6, d=256, ffn=2048, max_video_len=200,
                          app_num=20, mot_num=34, front_end_ch: Optional[Tuple[int, int, int]] = None,
                          text_tower: Optional[Tuple[int, int]] = None) -> Dict[str, Tuple[int, ...]]:
    """Names/shapes of every state_dict entry the hot path READS (subset of SURVEY.md §8b).  `front_end_ch` =
    (ResNet channels, Video-Swin channels, RoBERTa hidden) appends `input_proj`, `input_proj2` and
    `text_encoder.resizer` AFTER every other entry (so the hot-path weights of a seed do not depend on it).
    `text_tower` = (layers, vocab) appends the RoBERTa encoder `text_encoder.body.*` (hidden = front_end_ch[2]) after those."""
    s: Dict[str, Tuple[int, ...]] = {}

    def lin(name, o, i):
        s[name + ".weight"] = (o, i); s[name + ".bias"] = (o,)

    def ln(name, n=d):
        s[name + ".weight"] = (n,); s[name + ".bias"] = (n,)

    def mha(name):
        s[name + ".in_proj_weight"] = (3 * d, d); s[name + ".in_proj_bias"] = (3 * d,)
        lin(name + ".out_proj", d, d)

    for i in range(enc_layers):
        p = f"ground_encoder.encoder.spatial_layers.{i}."
        mha(p + "self_attn"); lin(p + "linear1", ffn, d); lin(p + "linear2", d, ffn); ln(p + "norm1"); ln(p + "norm2")
    ln("ground_encoder.encoder.norm")
    for c, vocab in (("t_temporal_clas", 1), ("s_temporal_clas", 1), ("t_spatial_clas", mot_num), ("s_spatial_clas", app_num)):
        for i in range(2):
            p = f"{c}.layer_ca.{i}."
            for n in ("query", "key", "value"):
                lin(p + "attention.self." + n, d, d)
            lin(p + "attention.output.dense", d, d); ln(p + "attention.output.LayerNorm")
            lin(p + "hidden_intermediate.dense", d, d); lin(p + "output.dense", d, d); ln(p + "output.LayerNorm")
        lin(c + ".head.transform.dense", d, d); ln(c + ".head.transform.LayerNorm")
        s[c + ".head.decoder.weight"] = (vocab, d); s[c + ".head.bias"] = (vocab,)
    g = "ground_decoder."
    ln(g + "pos_fc.0"); lin(g + "pos_fc.2", 4, d); ln(g + "pos_fc.4", 4)
    s[g + "time_embed.te"] = (max_video_len + 1, 1, d)
    for i in range(dec_layers):
        p = f"{g}time_decoder.layers.{i}."
        mha(p + "self_attn"); mha(p + "cross_attn_image"); lin(p + "linear1", ffn, d); lin(p + "linear2", d, ffn)
        ln(p + "norm1"); ln(p + "norm3"); ln(p + "norm4")
        p = f"{g}decoder.layers.{i}."
        for n in ("sa_qcontent_proj", "sa_qpos_proj", "sa_qtime_proj", "sa_kcontent_proj", "sa_kpos_proj",
                  "sa_ktime_proj", "sa_v_proj", "ca_qcontent_proj", "ca_kcontent_proj", "ca_kpos_proj",
                  "ca_v_proj", "ca_qpos_sine_proj"):
            lin(p + n, d, d)
        if i == 0:
            lin(p + "ca_qpos_proj", d, d)
        mha(p + "self_attn"); lin(p + "cross_attn.out_proj", d, d)
        lin(p + "linear1", ffn, d); lin(p + "linear2", d, ffn); ln(p + "norm1"); ln(p + "norm3"); ln(p + "norm4")
    ln(g + "time_decoder.norm")
    lin(g + "decoder.query_scale.layers.0", d, d); lin(g + "decoder.query_scale.layers.1", d, d)
    lin(g + "decoder.ref_point_head.layers.0", d, 2 * d); lin(g + "decoder.ref_point_head.layers.1", d, d)
    lin("bbox_embed.layers.0", d, d); lin("bbox_embed.layers.1", d, d); lin("bbox_embed.layers.2", 4, d)
    lin("temp_embed.layers.0", d, d); lin("temp_embed.layers.1", 2, d)
    lin("action_embed.layers.0", d, d); lin("action_embed.layers.1", 1, d)
    if front_end_ch is not None:
        cv, cd, ct = front_end_ch
        s["input_proj.weight"] = (d, cv, 1, 1); s["input_proj.bias"] = (d,)
        s["input_proj2.weight"] = (d, cd, 1, 1); s["input_proj2.bias"] = (d,)
        lin("text_encoder.resizer.fc", d, ct); ln("text_encoder.resizer.layer_norm")
    if text_tower is not None:
        layers, vocab = text_tower
        hd = front_end_ch[2]
        b = "text_encoder.body."
        s[b + "embeddings.word_embeddings.weight"] = (vocab, hd)
        s[b + "embeddings.position_embeddings.weight"] = (514, hd)
        s[b + "embeddings.token_type_embeddings.weight"] = (1, hd)
        ln(b + "embeddings.LayerNorm", hd)
        for i in range(layers):
            p = f"{b}encoder.layer.{i}."
            for n in ("query", "key", "value"):
                lin(p + "attention.self." + n, hd, hd)
            lin(p + "attention.output.dense", hd, hd); ln(p + "attention.output.LayerNorm", hd)
            lin(p + "intermediate.dense", 4 * hd, hd); lin(p + "output.dense", hd, 4 * hd); ln(p + "output.LayerNorm", hd)
    return s


def synth_state_dict(seed: int = 0, **kw) -> Dict[str, np.ndarray]:
    """Deterministic synthetic weights (numpy PCG64 — identical on every machine, no torch RNG) with the SCALES of
    the reference's own random init: encoder / decoder matrices ~ xavier_uniform (modal_encoder.py:36-39,
    query_decoder.py:71-74); classifier, bbox/temp/action-head matrices ~ nn.Linear default U(±1/sqrt(fan_in))
    (they are built outside / attached after the xavier reset, grounding_net.py:55-82).  Unlike the reference init,
    biases are non-zero U(±0.05) and LayerNorm gains are 1+U(±0.1) so that every parameter is exercised.
    `time_embed.te` is the real sine table.  Used instead of the reference's torch init so that golden fixtures
    stay small (the 36 M weights are regenerated from the seed, never stored)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: Dict[str, np.ndarray] = {}
    for name, shp in hot_path_param_shapes(**kw).items():
        if name.endswith("time_embed.te"):
            sd[name] = seq_embedding_sine(shp[0], shp[2])
        elif name.startswith("text_encoder.body.") and len(shp) >= 2:
            # transformers init: N(0, 0.02) for Linear / Embedding weights → uniform of the same std; a larger scale (x4) on the
            # Linear weights keeps the random-init tower away from the LayerNorm-only regime so that every matmul matters
            bound = 0.02 * math.sqrt(3.0) * (1.0 if "embeddings" in name else 4.0)
            sd[name] = rng.uniform(-bound, bound, size=shp).astype(F32)
        elif len(shp) >= 2:
            if name.startswith(("ground_encoder.", "ground_decoder.")):
                bound = math.sqrt(6.0 / (shp[0] + shp[1]))
            else:
                bound = 1.0 / math.sqrt(shp[1])
            sd[name] = rng.uniform(-bound, bound, size=shp).astype(F32)
        elif ("norm" in name.lower() and name.endswith(".weight")) or name.endswith(("pos_fc.0.weight", "pos_fc.4.weight")):
            sd[name] = (1.0 + rng.uniform(-0.1, 0.1, size=shp)).astype(F32)
        else:
            sd[name] = rng.uniform(-0.05, 0.05, size=shp).astype(F32)
    return sd


def synth_inputs(seed: int, T: int, H: int, W: int, L: int, d: int = 256):
    """Synthetic hot-path-boundary inputs: randn features, all-False masks, sine pos."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    vis = rng.standard_normal((T, d, H, W), dtype=F32)
    vid = rng.standard_normal((T, d, H, W), dtype=F32)
    text = rng.standard_normal((L, 1, d), dtype=F32)
    pos = sine_position_table(T, H, W)
    return vis, vid, pos, text


def synth_event_inputs(seed: int, T: int, H: int, W: int, L: int, amp: float = 2.0, d: int = 256, return_events: bool = False):
    """`synth_inputs` plus a temporal structure: two "events" (runs of frames that share an added channel pattern) and two
    single "spike" frames with patterns of their own, so that the per-frame relevance / actioness / start-end logits differ
    between frames by much more than the bf16 error of the path — the fixtures built on these inputs have partial frame
    selections in both decoder passes and wide decision margins (tests/golden/make_golden.py, `ev_*` cases)."""
    vis, vid, pos, text = synth_inputs(seed, T, H, W, L, d)
    rng = np.random.Generator(np.random.PCG64(9000 + seed))
    a = np.zeros((T, 4), F32)
    c1 = int(rng.integers(T // 8, max(T // 8 + 1, T // 2))); w1 = int(rng.integers(max(2, T // 8), max(3, T // 4)))
    a[c1:c1 + w1, 0] = 1
    c2 = int(rng.integers(T // 2, max(T // 2 + 1, T - T // 8))); w2 = int(rng.integers(max(1, T // 16), max(2, T // 6)))
    a[c2:c2 + w2, 1] = 1
    a[int(rng.integers(0, max(1, T // 2))), 2] = 2
    a[int(rng.integers(T // 2, T)), 3] = 2
    U = rng.standard_normal((4, d), dtype=F32)
    V = rng.standard_normal((4, d), dtype=F32)
    vis = (vis + F32(amp) * (a @ U)[:, :, None, None]).astype(F32)
    vid = (vid + F32(amp) * (a @ V)[:, :, None, None]).astype(F32)
    if return_events:
        return vis, vid, pos, text, a
    return vis, vid, pos, text


CALIB_PREFIX = "w:"


def apply_calibration(sd: Dict[str, np.ndarray], calib) -> Dict[str, np.ndarray]:
    """Override entries of a synthetic state dict (in place; returns it) with the arrays a fixture stores under
    `w:<state_dict key>`.  The "decisive" fixtures (tests/golden/make_golden.py, `ev_*`) re-design the last Linear layer of the
    two TemporalSampling heads, of `action_embed` and of `temp_embed` — 5 rows of 256 weights and their biases — from a run of
    the reference modules, so that the 0.45 / 0.5 thresholds and the start / end argmax fall into wide gaps of the reference's
    own per-frame scores (grounding_net.py:122-128,144-145; postprocessor.py:36-48)."""
    for k in (calib.files if hasattr(calib, "files") else calib.keys()):
        if k.startswith(CALIB_PREFIX):
            name = k[len(CALIB_PREFIX):]
            assert name in sd and sd[name].shape == calib[k].shape, name
            sd[name] = np.asarray(calib[k], F32)
    return sd


def synth_raw_inputs(seed: int, T: int, H: int, W: int, L: int, ch: Tuple[int, int, int] = (2048, 768, 768)):
    """Synthetic extractor outputs for the front end: a non-negative (post-ReLU, like ResNet layer 4) map, a
    normal Video-Swin map and normal RoBERTa hidden states."""
    rng = np.random.Generator(np.random.PCG64(5000 + seed))
    vis_raw = np.maximum(rng.standard_normal((T, ch[0], H, W), dtype=F32), 0)
    vid_raw = rng.standard_normal((T, ch[1], H, W), dtype=F32)
    text_raw = rng.standard_normal((L, ch[2]), dtype=F32)
    return vis_raw, vid_raw, text_raw


def synth_swin_stage(seed: int, dim: int = 768, heads: int = 24, window=(8, 7, 7), depth: int = 2, prefix: str = "vid.layers.3."):
    """Synthetic weights of the last Video-Swin stage (`BasicLayer`: `depth` SwinTransformerBlock3D, video_swin_transformer.py:
    176-275) under the reference's key names, at the scale of its init (trunc_normal 0.02 for Linear weights and the bias table —
    widened so that attention and bias matter against the random-init LayerNorm regime), non-zero biases, LayerNorm gains 1 ± 0.1."""
    rng = np.random.Generator(np.random.PCG64(11000 + seed))
    sd: Dict[str, np.ndarray] = {}
    u = lambda b, shp: rng.uniform(-b, b, size=shp).astype(F32)
    nb = (2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1)
    for i in range(depth):
        p = f"{prefix}blocks.{i}."
        for n in ("norm1", "norm2"):
            sd[p + n + ".weight"] = (1.0 + u(0.1, (dim,))).astype(F32)
            sd[p + n + ".bias"] = u(0.05, (dim,))
        sd[p + "attn.relative_position_bias_table"] = u(1.0, (nb, heads))
        for n, (o, k) in (("attn.qkv", (3 * dim, dim)), ("attn.proj", (dim, dim)), ("mlp.fc1", (4 * dim, dim)), ("mlp.fc2", (dim, 4 * dim))):
            sd[p + n + ".weight"] = u(2.5 / math.sqrt(k), (o, k))
            sd[p + n + ".bias"] = u(0.05, (o,))
    return sd


def synth_swin_backbone(seed: int, embed: int = 96, depths=(2, 2, 6, 2), heads=(3, 6, 12, 24), window=(8, 7, 7), prefix: str = "vid."):
    """Synthetic weights of the whole Video-Swin-T extractor as VSTGNet holds it (`self.vid`, VideoSwinTransformerBackbone:
    patch_embed.{proj,norm}, layers.{s}.blocks.{i}.*, downsamples.{s}.{norm,reduction}) — video_swin_transformer.py:626-664."""
    rng = np.random.Generator(np.random.PCG64(13000 + seed))
    u = lambda b, shp: rng.uniform(-b, b, size=shp).astype(F32)
    sd: Dict[str, np.ndarray] = {}
    sd[prefix + "patch_embed.proj.weight"] = u(1.0 / math.sqrt(48), (embed, 3, 1, 4, 4))
    sd[prefix + "patch_embed.proj.bias"] = u(0.05, (embed,))
    sd[prefix + "patch_embed.norm.weight"] = (1.0 + u(0.1, (embed,))).astype(F32)
    sd[prefix + "patch_embed.norm.bias"] = u(0.05, (embed,))
    for s, (dep, nh) in enumerate(zip(depths, heads)):
        dim = embed * 2 ** s
        stage = synth_swin_stage(seed * 10 + s, dim=dim, heads=nh, window=window, depth=dep, prefix=f"{prefix}layers.{s}.")
        sd.update(stage)
        if s + 1 < len(depths):
            sd[f"{prefix}downsamples.{s}.norm.weight"] = (1.0 + u(0.1, (4 * dim,))).astype(F32)
            sd[f"{prefix}downsamples.{s}.norm.bias"] = u(0.05, (4 * dim,))
            sd[f"{prefix}downsamples.{s}.reduction.weight"] = u(1.5 / math.sqrt(4 * dim), (2 * dim, 4 * dim))
    return sd


def synth_resnet101(seed: int, blocks=(3, 4, 23, 3), prefix: str = "vis_encoder.0.body."):
    """Synthetic weights of the ResNet101 extractor as VSTGNet holds it (`self.vis_encoder[0].body`: torchvision resnet101 with
    FrozenBatchNorm2d buffers, backbone.py:13-57,104-113): conv1 / bn1, layer{1-4}.{b}.conv{1,2,3} / bn{1,2,3} / downsample.{0,1}.
    He-scaled convolutions; the third BN of a block has a small gain so that 33 residual additions stay O(1)."""
    rng = np.random.Generator(np.random.PCG64(15000 + seed))
    sd: Dict[str, np.ndarray] = {}

    def conv(name, o, c, k):
        sd[prefix + name + ".weight"] = (rng.standard_normal((o, c, k, k)) * math.sqrt(2.0 / (c * k * k))).astype(F32)

    def bn(name, o, gain=1.0):
        sd[prefix + name + ".weight"] = (gain * rng.uniform(0.8, 1.2, size=(o,))).astype(F32)
        sd[prefix + name + ".bias"] = (0.05 * rng.standard_normal((o,))).astype(F32)
        sd[prefix + name + ".running_mean"] = (0.05 * rng.standard_normal((o,))).astype(F32)
        sd[prefix + name + ".running_var"] = rng.uniform(0.8, 1.2, size=(o,)).astype(F32)

    conv("conv1", 64, 3, 7); bn("bn1", 64)
    cin = 64
    for l, nb in enumerate(blocks):
        width = 64 << l
        for b in range(nb):
            q = f"layer{l + 1}.{b}."
            conv(q + "conv1", width, cin, 1); bn(q + "bn1", width)
            conv(q + "conv2", width, width, 3); bn(q + "bn2", width)
            conv(q + "conv3", 4 * width, width, 1); bn(q + "bn3", 4 * width, gain=0.3)
            if b == 0:
                conv(q + "downsample.0", 4 * width, cin, 1); bn(q + "downsample.1", 4 * width)
            cin = 4 * width
    return sd


def synth_text_ids(seed: int, B: int, L: int, vocab: int, pad_tail: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Token ids as RobertaTokenizer emits them: <s>=0 ... </s>=2, pad=1 on the last `pad_tail` positions of the odd rows.
    Returns (ids (B, L) int32, pad mask (B, L) bool, True = padded)."""
    rng = np.random.Generator(np.random.PCG64(7000 + seed))
    ids = rng.integers(3, vocab, size=(B, L)).astype(np.int32)
    ids[:, 0] = 0
    pad = np.zeros((B, L), bool)
    for b in range(B):
        n = L - (pad_tail if b % 2 == 1 else 0)
        ids[b, n - 1] = 2
        ids[b, n:] = 1
        pad[b, n:] = True
    return ids, pad


def synth_masks(masked: bool, T: int, H: int, W: int, L: int):
    """Padding masks for the masked golden case: right column on every frame, bottom row on the second
    half of the clip, two trailing text tokens (True = padded)."""
    vis_mask = np.zeros((T, H, W), bool)
    text_mask = np.zeros((1, L), bool)
    if masked:
        vis_mask[:, :, W - 1] = True
        vis_mask[T // 2:, H - 1, :] = True
        text_mask[0, L - 2:] = True
    return vis_mask, text_mask
