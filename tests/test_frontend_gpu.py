"""GPU parity of the front end: `vgqa_input_proj` (input_proj / input_proj2, the 1x1 Conv2d of
grounding_net.py:62,71,101,105 fused with the token-major re-layout) and the raw-input form of `vgqa_forward`
(input_proj + input_proj2 + text_encoder.resizer + the hot path), against the golden vectors of the reference's own
modules (tests/golden/fe_*.npz) and against the numpy oracle on seeded inputs.

Tolerances: the projection itself is compared (a) with an fp64 product of the SAME bf16-rounded operands → 2e-3 (fp32
accumulation order only) and (b) with the fp32 reference → 2e-2 (the north-star bf16 bar); the chained forward uses the
2e-2 bar of tests/test_parity_gpu.py with the reference's frame decisions forced."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import vgqa_oracle as O
from conftest import golden_path

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _load():
    from vgqa_b200 import _lib
    L = _lib.lib()
    L.vgqa_input_proj.restype = ctypes.c_int
    L.vgqa_input_proj.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return L, _lib


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float64).numpy()


def run_input_proj(F, C, H, W, S, tok0, pos_frames, seed, with_xp=True, with_x32=True):
    L, _lib = _load()
    P = H * W
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.maximum(rng.standard_normal((F, C, H, W), dtype=np.float32), 0)
    w = rng.uniform(-1, 1, (256, C)).astype(np.float32) / np.sqrt(C)
    b = rng.uniform(-0.05, 0.05, 256).astype(np.float32)
    pos = rng.standard_normal((pos_frames * S, 256), dtype=np.float32)
    dev = "cuda"
    tx = torch.from_numpy(x).to(dev)
    tw = torch.from_numpy(w).to(dev).to(torch.bfloat16).contiguous()
    tb = torch.from_numpy(b).to(dev)
    tpos = torch.from_numpy(pos).to(dev).to(torch.bfloat16).contiguous()
    sentinel = 768.0   # exactly representable in bf16
    X = torch.full((F * S, 256), sentinel, dtype=torch.bfloat16, device=dev)
    X32 = torch.full((F * S, 256), sentinel, dtype=torch.float32, device=dev)
    XP = torch.full((F * S, 256), sentinel, dtype=torch.bfloat16, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(L.vgqa_input_proj(p(tx), C, p(tw), p(tb), p(tpos) if with_xp else None, pos_frames, p(X),
                                 p(X32) if with_x32 else None, p(XP) if with_xp else None, F, S, tok0, P,
                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    # fp64 product of the bf16-rounded operands, token-major
    xr = bf16_round(x).reshape(F, C, P).transpose(0, 2, 1)           # [F, P, C]
    ref = xr @ bf16_round(w).T + b.astype(np.float64)                 # [F, P, 256]
    ref32 = (x.reshape(F, C, P).transpose(0, 2, 1).astype(np.float64) @ w.T.astype(np.float64) + b)
    X32n = X32.cpu().numpy().reshape(F, S, 256)
    Xn = X.float().cpu().numpy().reshape(F, S, 256)
    XPn = XP.float().cpu().numpy().reshape(F, S, 256)
    rows = slice(tok0, tok0 + P)
    if with_x32:
        assert float(np.abs(X32n[:, rows] - ref).max()) <= 2e-3
        assert float(np.abs(X32n[:, rows] - ref32).max()) <= TOL
    assert float(np.abs(Xn[:, rows] - ref).max()) <= 2e-3 + 2 ** -8 * float(np.abs(ref).max())
    if with_xp:
        posr = bf16_round(pos).reshape(pos_frames, S, 256)[:, rows]
        want = ref + (posr if pos_frames > 1 else posr[:1])
        assert float(np.abs(XPn[:, rows] - want).max()) <= 2e-3 + 2 ** -8 * float(np.abs(want).max())
    # rows outside [tok0, tok0 + P) of every frame are untouched
    other = np.ones(S, bool); other[rows] = False
    assert (Xn[:, other] == sentinel).all() and (X32n[:, other] == sentinel).all() and (XPn[:, other] == sentinel).all()
    if not with_x32:
        assert (X32n == sentinel).all()
    if not with_xp:
        assert (XPn == sentinel).all()


@pytest.mark.parametrize("F,C,H,W,L,second,pos_frames", [
    (1, 64, 1, 1, 1, False, 1),          # smallest: one token, one k-block
    (3, 128, 3, 4, 3, True, 1),          # ragged tile: 10 frames of 12 tokens per 128-row tile, 3 frames only
    (64, 2048, 7, 7, 20, False, 1),      # cfg-1/2 ResNet map: two 49-token frames per tile
    (64, 768, 7, 7, 20, True, 64),       # Video-Swin map written behind the text rows, per-frame pos rows
    (5, 768, 12, 12, 64, True, 1),       # 384 px: 144 tokens → two 128-row slices per frame
    (4, 2048, 14, 14, 20, False, 4),     # yaml 14x14: 196 tokens
    (301, 192, 7, 7, 20, True, 1),       # more tiles than SMs, odd frame count (half-empty last tile)
])
def test_input_proj_kernel(F, C, H, W, L, second, pos_frames):
    P = H * W
    S = 2 * P + L
    run_input_proj(F, C, H, W, S, (P + L) if second else 0, pos_frames, seed=F * 1000 + C)


def test_input_proj_optional_outputs_and_errors():
    run_input_proj(7, 128, 4, 4, 40, 0, 1, seed=3, with_xp=False, with_x32=False)
    L, _lib = _load()
    z = torch.zeros(16, device="cuda")
    p = ctypes.c_void_p(z.data_ptr())
    assert L.vgqa_input_proj(p, 100, p, p, None, 1, p, None, None, 1, 3, 0, 1, None) != 0      # C % 64 != 0
    assert b"multiple of 64" in L.vgqa_last_error()
    assert L.vgqa_input_proj(p, 64, p, p, None, 1, p, None, None, 1, 3, 2, 2, None) != 0       # tok0 + P > S
    assert L.vgqa_input_proj(p, 64, p, p, None, 1, p, None, p, 1, 3, 0, 1, None) != 0          # XP without pos


FE_CASES = ["fe_tiny_T3_3x4_L3", "fe_cfg1_T32_7x7_L20_s0", "fe_yaml_T4_14x14_L20_s1"]


def _engine_and_inputs(name, clips=1):
    from vgqa_b200.engine import GroundingEngine
    g = np.load(golden_path(name))
    T, H, W, L, seed = (int(g[k]) for k in ("T", "H", "W", "L", "seed"))
    ch = tuple(int(x) for x in g["front_end_ch"])
    sd = O.synth_state_dict(seed, front_end_ch=ch)
    eng = GroundingEngine(sd, max_clips=clips, max_frames=max(T, 4), max_hw=H * W, max_text=L)
    raw = O.synth_raw_inputs(seed, T, H, W, L, ch)
    pos = O.position_embedding_sine(np.zeros((T, H, W), bool))
    rep = lambda a: torch.from_numpy(np.ascontiguousarray(np.stack([a] * clips))).cuda()
    return g, sd, eng, raw, pos, rep


@pytest.mark.parametrize("name", FE_CASES)
def test_raw_forward_matches_reference_chain(name):
    g, sd, eng, (vis_raw, vid_raw, text_raw), pos, rep = _engine_and_inputs(name)
    T = int(g["T"])
    w1 = np.zeros(T, np.float32); w1[g["choose_pass1"]] = 1
    w2 = np.zeros(T, np.float32); w2[g["choose_pass2"]] = 1
    sizes = torch.tensor([[float(g["ori_size"][0]), float(g["ori_size"][1])]], device="cuda")
    o = eng.forward(rep(vis_raw), rep(vid_raw), rep(text_raw), torch.from_numpy(pos[:1].copy()).cuda(), ori_sizes_hw=sizes,
                    force_choose1=rep(w1), force_choose2=rep(w2), raw=True)
    torch.cuda.synchronize()
    o = {k: v.cpu().numpy() for k, v in o.items()}
    cmp = {"pred_boxes": g["pred_boxes"], "pred_sted": g["pred_sted"][0], "pred_actioness": g["pred_actioness"][0, :, 0],
           "logits_f_m": g["logits_f_m"], "logits_f_a": g["logits_f_a"], "logits_r_a": g["logits_r_a"][0],
           "logits_r_m": g["logits_r_m"][0], "att_sequences": g["att_sequences"][0], "actioness_pass1": g["actioness_pass1"]}
    worst = {k: float(np.abs(o[k][0].reshape(ref.shape) - ref).max()) for k, ref in cmp.items()}
    worst["frames_cls"] = float(np.abs(o["frames_cls"] - g["frames_cls"]).max())
    bad = {k: v for k, v in worst.items() if not v <= TOL}
    assert not bad, f"{name}: max-abs errors over {TOL}: {bad} (all: {worst})"
    np.testing.assert_allclose(o["boxes_px"][0], g["post_boxes"], atol=TOL * 640)
    eng.close()


def test_raw_forward_equals_projected_forward():
    """Same engine, same clip: raw inputs through the fused front end vs the oracle's fp32 front end fed to the
    projected-input entry.  Differences are only the bf16 rounding of the front-end operands."""
    g, sd, eng, raw, pos, rep = _engine_and_inputs("fe_cfg1_T32_7x7_L20_s0", clips=2)
    vis, vid, text = O.front_end(sd, *raw)
    tpos = torch.from_numpy(pos[:1].copy()).cuda()
    want = ["pred_boxes", "pred_sted", "logits_f_m", "logits_f_a", "frames_cls"]
    a = eng.forward(rep(raw[0]), rep(raw[1]), rep(raw[2]), tpos, raw=True, want=want)
    b = eng.forward(rep(vis), rep(vid), rep(text[:, 0]), tpos, want=want)
    torch.cuda.synchronize()
    for k in want:
        x, y = a[k].cpu().numpy(), b[k].cpu().numpy()
        x2 = x.reshape(2, -1)
        np.testing.assert_array_equal(x2[0], x2[1])        # both clips of the batch are the same clip
        assert float(np.abs(x - y).max()) <= TOL, k
    # host-buffer entry with raw inputs (pinned staging of the raw maps)
    h = eng.forward_host(rep(raw[0]).cpu(), rep(raw[1]).cpu(), rep(raw[2]).cpu(), tpos.cpu(), raw=True, want=want)
    for k in want:
        np.testing.assert_allclose(h[k].numpy(), a[k].cpu().numpy(), atol=1e-6)
    eng.close()


def test_raw_inputs_need_their_weights():
    from vgqa_b200.engine import GroundingEngine
    eng = GroundingEngine(O.synth_state_dict(0), max_clips=1, max_frames=8, max_hw=9, max_text=8)   # no front-end weights
    z = lambda *s: torch.zeros(*s, device="cuda")
    with pytest.raises(RuntimeError, match="input_proj"):
        eng.forward(z(1, 4, 128, 3, 3), z(1, 4, 64, 3, 3), z(1, 4, 64), z(1, 256, 3, 3), raw=True)
    eng.close()
    sd = O.synth_state_dict(0, front_end_ch=(128, 64, 64))
    eng = GroundingEngine(sd, max_clips=1, max_frames=8, max_hw=9, max_text=8)
    with pytest.raises(RuntimeError, match="vis_raw_ch"):
        eng.forward(z(1, 4, 192, 3, 3), z(1, 4, 64, 3, 3), z(1, 4, 64), z(1, 256, 3, 3), raw=True)
    eng.close()


def test_vstgnet_dropin_and_predictor_with_fused_front_end():
    """The Python seam with raw extractor outputs: B200VSTGNet skips the torch input_proj modules (they are never called) and
    GroundingPredictor(raw_inputs=True) serves the even/odd passes from raw maps; both against the reference golden."""
    from make_golden import make_cfg
    from vgqa_b200 import modules as M
    from vgqa_b200.predict import GroundingPredictor
    g = np.load(golden_path("fe_cfg1_T32_7x7_L20_s0"))
    T, H, W, L, seed = (int(g[k]) for k in ("T", "H", "W", "L", "seed"))
    ch = tuple(int(x) for x in g["front_end_ch"])
    sd = O.synth_state_dict(seed, front_end_ch=ch)
    vis_raw, vid_raw, text_raw = (torch.from_numpy(a).cuda() for a in O.synth_raw_inputs(seed, T, H, W, L, ch))
    pos = torch.from_numpy(O.position_embedding_sine(np.zeros((T, H, W), bool))).cuda()

    class Boom(torch.nn.Module):
        def forward(self, x):
            raise AssertionError("the torch input_proj must not run when the front end is fused")

    vis_encoder = lambda videos: (M.NestedTensor(vis_raw, videos.mask[:, :H, :W], videos.durations), pos)
    vid_model = lambda tensors, n: {"3": vid_raw}
    text_encoder = lambda texts, device: ((torch.zeros(1, L, dtype=torch.bool, device=device), None, text_raw[:, None, :]), None)
    model = M.B200VSTGNet(make_cfg(), vis_encoder, vid_model, text_encoder, Boom(), Boom(), sd, verb_label2={"0": {"sub": ""}},
                          max_frames=64, max_hw=49, max_text=20).eval()
    assert model.fused_front_end
    videos = M.NestedTensor(torch.zeros(T, 3, 224, 224, device="cuda"), torch.zeros(T, 224, 224, dtype=torch.bool, device="cuda"), [T])
    out = model(videos, ["a person jumping"], [{"item_id": 0, "actioness": torch.ones(T, device="cuda")}])
    for k in ("pred_boxes", "pred_sted", "logits_f_m", "logits_f_a", "att_sequences"):
        assert float(np.abs(out[k].cpu().numpy() - g[k]).max()) <= TOL, k
    # predictor: 2T sampled frames whose even pass is the golden clip
    pred = GroundingPredictor(sd, sample_num=T, max_hw=H * W, max_text=L, raw_inputs=True)
    il = lambda a: torch.stack([a, a.flip(0)], 1).reshape(2 * T, *a.shape[1:]).contiguous()   # even frames = the clip
    fids = list(range(2 * T))
    res = pred.predict(il(vis_raw), il(vid_raw), text_raw, pos[:1], fids, (360, 640), fps=25.0)
    even = {t["frame"]: t["bbox"] for t in res["tube"] if t["frame"] % 2 == 0}
    got = np.asarray([even[f] for f in range(0, 2 * T, 2)])
    np.testing.assert_allclose(got, g["post_boxes"], atol=TOL * 640)
    pred.close()


@pytest.mark.parametrize("F,C,H,W,L,second,pos_frames", [(1, 64, 1, 1, 1, False, 1), (64, 2048, 7, 7, 20, False, 1),
                                                          (37, 768, 7, 7, 20, True, 37), (5, 768, 12, 12, 64, True, 1),
                                                          (301, 192, 7, 7, 20, True, 1)])
def test_input_proj_nhwc_bf16_kernel(F, C, H, W, L, second, pos_frames):
    """Channels-last bf16 maps (raw_layout = 1): the rows are the TMA-fed A operand; vs an fp64 product of the same bf16 values."""
    Lb, _lib = _load()
    Lb.vgqa_input_proj_nhwc.restype = ctypes.c_int
    Lb.vgqa_input_proj_nhwc.argtypes = Lb.vgqa_input_proj.argtypes
    P = H * W
    S = 2 * P + L
    tok0 = (P + L) if second else 0
    g = torch.Generator(device="cuda").manual_seed(F * 31 + C)
    x = torch.randn(F, P, C, device="cuda", generator=g).relu_().to(torch.bfloat16)
    w = (torch.randn(256, C, device="cuda", generator=g) / C ** 0.5).to(torch.bfloat16)
    b = torch.randn(256, device="cuda", generator=g) * 0.05
    pos = torch.randn(pos_frames * S, 256, device="cuda", generator=g).to(torch.bfloat16)
    X = torch.full((F * S, 256), 768.0, dtype=torch.bfloat16, device="cuda")
    X32 = torch.full((F * S, 256), 768.0, dtype=torch.float32, device="cuda")
    XP = torch.full((F * S, 256), 768.0, dtype=torch.bfloat16, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(Lb.vgqa_input_proj_nhwc(p(x), C, p(w), p(b), p(pos), pos_frames, p(X), p(X32), p(XP), F, S, tok0, P,
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    ref = (x.double() @ w.double().T + b.double()).cpu().numpy()                    # [F, P, 256]
    rows = slice(tok0, tok0 + P)
    X32n = X32.cpu().numpy().reshape(F, S, 256)
    assert float(np.abs(X32n[:, rows] - ref).max()) <= 2e-3
    posr = pos.double().cpu().numpy().reshape(pos_frames, S, 256)[:, rows]
    want = ref + (posr if pos_frames > 1 else posr[:1])
    XPn = XP.float().cpu().numpy().reshape(F, S, 256)
    assert float(np.abs(XPn[:, rows] - want).max()) <= 2e-3 + 2 ** -8 * float(np.abs(want).max())
    other = np.ones(S, bool); other[rows] = False
    assert (X32n[:, other] == 768.0).all() and (XPn[:, other] == 768.0).all()
    assert (X.float().cpu().numpy().reshape(F, S, 256)[:, other] == 768.0).all()


def test_forward_channels_last_bf16_equals_nchw_fp32():
    """vgqa_forward with channels-last bf16 maps vs the same (bf16-rounded) values given as the reference's NCHW fp32 maps."""
    g, sd, eng, raw, pos, rep = _engine_and_inputs("fe_cfg1_T32_7x7_L20_s0", clips=2)
    tpos = torch.from_numpy(pos[:1].copy()).cuda()
    want = ["pred_boxes", "pred_sted", "logits_f_m", "frames_cls"]
    vis16, vid16 = rep(raw[0]).to(torch.bfloat16), rep(raw[1]).to(torch.bfloat16)       # [2, T, C, H, W]
    a = eng.forward(vis16.float(), vid16.float(), rep(raw[2]), tpos, raw=True, want=want)
    b = eng.forward(vis16.permute(0, 1, 3, 4, 2).contiguous(), vid16.permute(0, 1, 3, 4, 2).contiguous(), rep(raw[2]), tpos,
                    raw=True, want=want)
    torch.cuda.synchronize()
    for k in want:
        assert float((a[k] - b[k]).abs().max()) <= 5e-3, k
    h = eng.forward_host(vis16.permute(0, 1, 3, 4, 2).contiguous().cpu(), vid16.permute(0, 1, 3, 4, 2).contiguous().cpu(),
                         rep(raw[2]).cpu(), tpos.cpu(), raw=True, want=want)
    for k in want:
        np.testing.assert_allclose(h[k].numpy(), b[k].cpu().numpy(), atol=1e-6)
    eng.close()


def test_hot_path_module_takes_bf16_channels_last_maps_zero_copy():
    """HotPath.forward(raw=True) with bf16 channels_last backbone outputs: same result as the fp32 NCHW form of the same values."""
    from make_golden import make_cfg
    from vgqa_b200 import modules as M
    g = np.load(golden_path("fe_cfg1_T32_7x7_L20_s0"))
    T, H, W, L, seed = (int(g[k]) for k in ("T", "H", "W", "L", "seed"))
    ch = tuple(int(x) for x in g["front_end_ch"])
    sd = O.synth_state_dict(seed, front_end_ch=ch)
    vis_raw, vid_raw, text_raw = (torch.from_numpy(a).cuda() for a in O.synth_raw_inputs(seed, T, H, W, L, ch))
    pos = torch.from_numpy(O.position_embedding_sine(np.zeros((T, H, W), bool))).cuda()
    hot = M.build_hot_path(make_cfg(), sd, max_frames=64, max_hw=49, max_text=20)
    vm = torch.zeros(T, H, W, dtype=torch.bool, device="cuda")
    tm = torch.zeros(1, L, dtype=torch.bool, device="cuda")
    v16 = vis_raw.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    d16 = vid_raw.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    assert v16.permute(0, 2, 3, 1).is_contiguous()                 # the view the module hands to the library
    a = hot(v16, vm, pos, tm, text_raw[:, None, :], d16, raw=True)
    b = hot(v16.float().contiguous(), vm, pos, tm, text_raw[:, None, :], d16.float().contiguous(), raw=True)
    for k in ("pred_boxes", "pred_sted", "logits_f_m", "att_sequences"):
        assert float((a[k] - b[k]).abs().max()) <= 5e-3, k
        assert float(np.abs(a[k].cpu().numpy() - g[k]).max()) <= 3e-2, k       # bf16-rounded extractor maps vs the fp32 golden
