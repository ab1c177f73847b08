"""GPU parity of the whole hot path (C-ABI vgqa_forward) against the golden vectors produced by the reference's
own PyTorch modules (tests/golden/*.npz) and against the numpy oracle on the same seeded inputs.

Bar (BASELINE.json north_star): start/end argmax frames and tube frame indices identical; boxes and logits within
2e-2 max-abs under bf16.

Two families of fixtures (tests/golden/make_golden.py):
  * `ev_*` — "decisive" cases: event-structured inputs and re-designed last-layer heads, so that the reference's own decisions
    are PARTIAL in both decoder passes (0 < K1 < T, 0 < K2 < T) and sit far from their thresholds (|att - 0.45| >= 0.02,
    |sigmoid(actioness) - 0.5| >= 0.02, top-2 gap of the start/end score map >= 0.05).  For these the free-running forward
    (nothing forced) must reproduce both frame selections, the start/end argmax, the tube frame ids and every continuous
    output — unconditional asserts.
  * the iid-random cases of round 1 (every frame chosen, or a fallback): continuous outputs with the reference's decisions
    forced through `force_choose1/2`; their free-running decisions are compared where the stored margin allows and the
    test SKIPS (never passes silently) where it does not."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import vgqa_oracle as O
from conftest import golden_path

pytestmark = pytest.mark.gpu

TOL = 2e-2
# frames_cls is not an output of the reference forward but an internal activation of the encoder (token mean of LayerNorm rows, |x| up
# to 4 — one bf16 ulp there is 1.6e-2); it is compared as a tripwire with a bound of its own
TOL_INTERNAL = 3e-2
CASES = ["tiny_T3_3x4_L3", "ragged_T6_4x5_L7_masked", "cfg1_T32_7x7_L20_s0", "cfg1_T32_7x7_L20_s1",
         "cfg2_T64_7x7_L20_s0", "cfg2_T64_7x7_L20_s2", "yaml_T16_14x14_L20_s0", "cfg4_T256_7x7_L20_s0",
         "cfg5_T128_12x12_L64_s0"]

from conftest import GOLDEN  # noqa: E402

EV_CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "ev_*.npz")))
MIN_THETA, MIN_ACT, MIN_TOP2 = 0.02, 0.02, 0.05   # the bounds make_golden.py builds the ev_* fixtures to

_engines = {}


def engine_for(g, name, T, P, L):
    """One engine per (weights, capacity).  Fixtures with `w:*` overrides (ev_*) get their own weights."""
    from vgqa_b200.engine import GroundingEngine
    seed, max_len = int(g["seed"]), int(g["max_video_len"])
    calibrated = any(k.startswith(O.CALIB_PREFIX) for k in g.files)
    key = (seed, max_len, name if calibrated else None)
    cap = _engines.get(key)
    if cap is None or cap[1] < T or cap[2] < P or cap[3] < L:
        if cap is not None:
            cap[0].close()
        for k in [k for k in _engines if k[2] is not None and k != key]:   # calibrated engines are single-use: free them
            _engines.pop(k)[0].close()
        sd = O.apply_calibration(O.synth_state_dict(seed, max_video_len=max_len), g)
        eng = GroundingEngine(sd, max_clips=2, max_frames=max(T, 64), max_hw=max(P, 49), max_text=max(L, 20),
                              max_video_len=max_len)
        cap = (eng, max(T, 64), max(P, 49), max(L, 20))
        _engines[key] = cap
    return cap[0]


def case_inputs(g):
    T, H, W, L, seed = (int(g[k]) for k in ("T", "H", "W", "L", "seed"))
    amp = float(g["event_amp"]) if "event_amp" in g.files else 0.0
    if amp > 0:
        return O.synth_event_inputs(seed, T, H, W, L, amp=amp)
    return O.synth_inputs(seed, T, H, W, L)


def run_case(name, force, clips=1):
    g = np.load(golden_path(name))
    T, H, W, L, seed = (int(g[k]) for k in ("T", "H", "W", "L", "seed"))
    masked = bool(g["masked"])
    eng = engine_for(g, name, T, H * W, L)
    vis, vid, _, text = case_inputs(g)
    itr = int(g["iteration_rate"]) if "iteration_rate" in g.files else -1
    vm, tm = O.synth_masks(masked, T, H, W, L)
    pos = O.position_embedding_sine(vm)
    dev = "cuda"
    rep = lambda a: torch.from_numpy(np.ascontiguousarray(np.stack([a] * clips))).to(dev)
    tvis, tvid = rep(vis), rep(vid)
    ttext = rep(text[:, 0, :])
    if masked:
        tpos = torch.from_numpy(np.ascontiguousarray(np.concatenate([pos] * clips))).to(dev)
        tvm = torch.from_numpy(np.concatenate([vm.reshape(T, -1)] * clips).astype(np.uint8)).to(dev)
        ttm = torch.from_numpy(np.concatenate([tm] * clips).astype(np.uint8)).to(dev)
    else:
        tpos, tvm, ttm = torch.from_numpy(pos[:1].copy()).to(dev), None, None
    sizes = torch.tensor([[float(g["ori_size"][0]), float(g["ori_size"][1])]] * clips, device=dev)
    f1 = f2 = None
    if force:
        w1 = np.zeros(T, np.float32); w1[g["choose_pass1"]] = 1
        w2 = np.zeros(T, np.float32); w2[g["choose_pass2"]] = 1
        f1, f2 = rep(w1), rep(w2)
    outs = eng.forward(tvis, tvid, ttext, tpos, vis_mask=tvm, text_mask=ttm, ori_sizes_hw=sizes, force_choose1=f1,
                       force_choose2=f2, iteration_rate=itr, want=None)
    torch.cuda.synchronize()
    return g, {k: v.cpu().numpy() for k, v in outs.items()}


def continuous_errors(g, o):
    cmp = {"pred_boxes": g["pred_boxes"], "pred_sted": g["pred_sted"][0], "pred_actioness": g["pred_actioness"][0, :, 0],
           "logits_f_m": g["logits_f_m"], "logits_f_a": g["logits_f_a"], "logits_r_a": g["logits_r_a"][0],
           "logits_r_m": g["logits_r_m"][0], "att_sequences": g["att_sequences"][0],
           "aux_boxes": g["aux_boxes"], "aux_sted": g["aux_sted"][:, 0], "aux_actioness": g["aux_actioness"][:, 0, :, 0],
           "frames_cls": g["frames_cls"], "actioness_pass1": g["actioness_pass1"]}
    if "iteration_rate" in g.files and int(g["iteration_rate"]) >= 0:
        cmp.pop("actioness_pass1")        # single-pass forward (grounding_net.py:143): the re-selection score is never computed
    worst = {}
    for k, ref in cmp.items():
        got = o[k][0] if k in ("pred_boxes", "pred_sted", "pred_actioness", "logits_f_m", "logits_f_a", "logits_r_a",
                                "logits_r_m", "att_sequences", "actioness_pass1") else o[k]
        if k.startswith("aux_"):
            got = o[k][:, 0]
        worst[k] = float(np.abs(got.reshape(ref.shape) - ref).max())
    return worst


@pytest.mark.parametrize("name", CASES + EV_CASES)
def test_continuous_outputs_with_reference_decisions(name):
    g, o = run_case(name, force=True)
    T = int(g["T"])
    worst = continuous_errors(g, o)
    bad = {k: v for k, v in worst.items() if not v <= (TOL_INTERNAL if k == "frames_cls" else TOL)}
    assert not bad, f"{name}: max-abs errors over {TOL}: {bad} (all: {worst})"
    # PostProcess (postprocessor.py:36-48): the (start,end) argmax must be the reference's whenever the reference's
    # top-2 gap is resolvable (> 4x the measured logit error of this run); it must always be near-optimal under the
    # reference's own scores.
    sted_ref = g["pred_sted"][0]
    score = O.log_softmax(sted_ref[:, 0], 0)[:, None] + O.log_softmax(sted_ref[:, 1], 0)[None, :]
    score = np.where(np.triu(np.ones((T, T), bool), 1), score, -np.inf)
    s, e = (int(x) for x in o["sted_idx"][0])
    assert s < e
    assert score[s, e] >= score.max() - 4 * max(worst["pred_sted"], 1e-3)
    if float(g["margin_sted_top2"]) >= MIN_TOP2 or float(g["margin_sted_top2"]) > 4 * worst["pred_sted"]:
        fid = g["frame_ids"]
        assert [int(fid[s]), int(fid[e]) + 1] == g["post_sted"][0].tolist()
    np.testing.assert_allclose(o["boxes_px"][0], g["post_boxes"], atol=TOL * 640)


def _ref_sel(g, key):
    w = np.zeros(int(g["T"]))
    w[g[key]] = 1
    return w


def test_decisive_fixtures_clear_the_bounds():
    """At least 8 of the 9 two-pass `ev_*` fixtures (+ the single-pass one) are partial in both passes with wide margins."""
    two_pass = [n for n in EV_CASES if int(np.load(golden_path(n))["iteration_rate"]) < 0]
    assert len(two_pass) >= 8 and len(EV_CASES) > len(two_pass), EV_CASES
    clear = 0
    for n in two_pass:
        g = np.load(golden_path(n))
        T = int(g["T"])
        clear += (0 < len(g["choose_pass1"]) < T and 0 < len(g["choose_pass2"]) < T and float(g["margin_theta"]) >= MIN_THETA and
                  float(g["margin_act"]) >= MIN_ACT and float(g["margin_sted_top2"]) >= MIN_TOP2)
    assert clear >= 8, f"only {clear} of {len(two_pass)} decisive fixtures clear the margin bounds"


@pytest.mark.parametrize("name", EV_CASES)
def test_free_running_decisions_identical(name):
    """Nothing forced: both frame selections, the start/end argmax, the tube frame ids and all continuous outputs must be
    the reference's (grounding_net.py:125-128,143-163; postprocessor.py:36-48).  Unconditional."""
    g, o = run_case(name, force=False)
    T = int(g["T"])
    two_pass = int(g["iteration_rate"]) < 0
    assert float(g["margin_theta"]) >= MIN_THETA and float(g["margin_sted_top2"]) >= MIN_TOP2, "fixture is not decisive"
    assert 0 < len(g["choose_pass1"]) < T
    np.testing.assert_array_equal(o["choose1"][0], _ref_sel(g, "choose_pass1"), "pass-1 frame selection")
    if two_pass:
        assert float(g["margin_act"]) >= MIN_ACT and 0 < len(g["choose_pass2"]) < T
        np.testing.assert_array_equal(o["choose2"][0], _ref_sel(g, "choose_pass2"), "pass-2 frame selection")
    worst = continuous_errors(g, o)
    bad = {k: v for k, v in worst.items() if not v <= (TOL_INTERNAL if k == "frames_cls" else TOL)}
    assert not bad, f"{name} (free-running): max-abs errors over {TOL}: {bad} (all: {worst})"
    s, e = (int(x) for x in o["sted_idx"][0])
    fid = g["frame_ids"]
    assert [int(fid[s]), int(fid[e]) + 1] == g["post_sted"][0].tolist(), "start/end argmax frames"
    # tube frame ids = the sampled frame ids inside [start, end) (evaluator.py:78-92 keeps one box per sampled frame)
    tube = [int(f) for f in fid if int(fid[s]) <= f < int(fid[e]) + 1]
    ref_tube = [int(f) for f in fid if g["post_sted"][0][0] <= f < g["post_sted"][0][1]]
    assert tube == ref_tube
    np.testing.assert_allclose(o["boxes_px"][0], g["post_boxes"], atol=TOL * 640)


@pytest.mark.parametrize("name", CASES)
def test_free_running_pass1_selection(name):
    g, o = run_case(name, force=False)
    att = g["att_sequences"][0]
    if not (att > 0.45).any():
        # the reference fell back to "every frame" (grounding_net.py:128): so must the library
        assert (o["choose1"][0] == 1).all()
        return
    safe = np.abs(att - 0.45) > 1e-2
    if not safe.any():
        pytest.skip(f"every frame is within 1e-2 of theta (margin {float(g['margin_theta']):.4f})")
    ref1 = _ref_sel(g, "choose_pass1")
    assert (o["choose1"][0][safe] == ref1[safe]).all(), "pass-1 frame selection differs outside the margin"


@pytest.mark.parametrize("name", CASES)
def test_free_running_pass2_selection_and_outputs(name):
    g, o = run_case(name, force=False)
    ref1, ref2 = _ref_sel(g, "choose_pass1"), _ref_sel(g, "choose_pass2")
    if not (o["choose1"][0] == ref1).all():
        pytest.skip(f"pass-1 selection differs inside the theta margin ({float(g['margin_theta']):.4f}): pass 2 is not comparable")
    safe = np.abs(g["actioness_pass1"] - 0.5) > 1e-2
    assert (o["choose2"][0][safe] == ref2[safe]).all(), "pass-2 frame selection differs outside the margin"
    if not (o["choose2"][0] == ref2).all():
        pytest.skip(f"pass-2 selection differs inside the 0.5 margin ({float(g['margin_act']):.4f})")
    assert float(np.abs(o["pred_boxes"][0] - g["pred_boxes"]).max()) <= TOL
    assert float(np.abs(o["pred_sted"][0] - g["pred_sted"][0]).max()) <= TOL


@pytest.mark.parametrize("name", ["ev_cfg1_T32_7x7_L20_s0", "ev_cfg2_T64_7x7_L20_a_s0"])
def test_bf16_channels_last_features(name):
    """vgqa_inputs.feat_layout = 1: the projected maps handed over as channels-last bf16 [clips, T, H, W, 256] (device and host
    entry points) give the reference's decisions and outputs within the same 2e-2 bar."""
    g = np.load(golden_path(name))
    T, H, W, L = (int(g[k]) for k in ("T", "H", "W", "L"))
    eng = engine_for(g, name, T, H * W, L)
    vis, vid, _, text = case_inputs(g)
    cl = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(0, 2, 3, 1))[None]).to(torch.bfloat16)
    sizes = torch.tensor([[float(g["ori_size"][0]), float(g["ori_size"][1])]])
    ttext = torch.from_numpy(np.ascontiguousarray(text[None, :, 0, :]))
    o = eng.forward(cl(vis).cuda(), cl(vid).cuda(), ttext.cuda(), None, ori_sizes_hw=sizes.cuda())
    torch.cuda.synchronize()
    o = {k: v.cpu().numpy() for k, v in o.items()}
    oh = eng.forward_host(cl(vis).pin_memory(), cl(vid).pin_memory(), ttext.pin_memory(), None, ori_sizes_hw=sizes.pin_memory())
    oh = {k: v.numpy() for k, v in oh.items()}
    for out in (o, oh):
        np.testing.assert_array_equal(out["choose1"][0], _ref_sel(g, "choose_pass1"))
        np.testing.assert_array_equal(out["choose2"][0], _ref_sel(g, "choose_pass2"))
        worst = continuous_errors(g, out)
        bad = {k: v for k, v in worst.items() if not v <= (TOL_INTERNAL if k == "frames_cls" else TOL)}
        assert not bad, f"{name} (bf16 channels-last features): {bad} (all: {worst})"
        s_, e_ = (int(x) for x in out["sted_idx"][0])
        fid = g["frame_ids"]
        assert [int(fid[s_]), int(fid[e_]) + 1] == g["post_sted"][0].tolist()
    for k in ("pred_boxes", "pred_sted", "logits_f_m"):
        np.testing.assert_allclose(oh[k], o[k], atol=1e-6, err_msg=k)     # host and device entry points run the same kernels


def test_fused_decoder_tail_chain(monkeypatch):
    """VGQA_CHAIN=1: the frame-local tail of every decoder layer ([vo + LN3] → FFN + LN4 → norm / next in-projection / bbox_embed)
    runs as ONE fused row-tile GEMM chain (csrc/chain.cu).  Same decisions and outputs as the reference, and fewer launches."""
    from vgqa_b200.engine import GroundingEngine
    name = "ev_cfg2_T64_7x7_L20_a_s0"
    g = np.load(golden_path(name))
    T, H, W, L, seed = (int(g[k]) for k in ("T", "H", "W", "L", "seed"))
    sd = O.apply_calibration(O.synth_state_dict(seed, max_video_len=int(g["max_video_len"])), g)
    vis, vid, _, text = case_inputs(g)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    sizes = torch.tensor([[float(g["ori_size"][0]), float(g["ori_size"][1])]] * 3, device="cuda")
    launches = {}
    outs = {}
    monkeypatch.setenv("VGQA_CHAIN_HEAD", "0")
    for flag in ("0", "1", "head"):
        monkeypatch.setenv("VGQA_CHAIN", "1" if flag == "1" else "0")          # read when the context is created
        monkeypatch.setenv("VGQA_CHAIN_HEAD", "1" if flag == "head" else "0")  # the light chain only: bbox_embed → sine → ref_point_head
        eng = GroundingEngine(sd, max_clips=3, max_frames=T, max_hw=H * W, max_text=L)
        o = eng.forward(t(np.stack([vis] * 3)), t(np.stack([vid] * 3)), t(np.stack([text[:, 0]] * 3)), None, ori_sizes_hw=sizes)
        torch.cuda.synchronize()
        outs[flag] = {k: v.cpu().numpy() for k, v in o.items()}
        launches[flag] = eng.last_launch_count
        eng.close()
    assert launches["1"] < launches["0"] - 100 and launches["head"] < launches["0"] - 50, launches   # 192 rows = two row tiles (one ragged)
    for flag in ("1", "head"):
        o = dict(outs[flag], frames_cls=outs[flag]["frames_cls"][:T])     # continuous_errors looks at clip 0
        for b in range(3):
            np.testing.assert_array_equal(o["choose1"][b], _ref_sel(g, "choose_pass1"))
            np.testing.assert_array_equal(o["choose2"][b], _ref_sel(g, "choose_pass2"))
        worst = continuous_errors(g, o)
        bad = {k: v for k, v in worst.items() if not v <= (TOL_INTERNAL if k == "frames_cls" else TOL)}
        assert not bad, f"chain {flag}: {bad} (all: {worst})"
        for k in ("pred_boxes", "pred_sted", "pred_actioness", "aux_boxes"):
            np.testing.assert_allclose(outs[flag][k], outs["0"][k], atol=5e-3, err_msg=k)   # same math, different summation order


def test_batch_of_clips_matches_single():
    g, o1 = run_case("cfg1_T32_7x7_L20_s1", force=False, clips=1)
    _, o2 = run_case("cfg1_T32_7x7_L20_s1", force=False, clips=2)
    for k in ("pred_boxes", "pred_sted", "logits_f_m", "logits_r_a"):
        np.testing.assert_array_equal(o2[k][0], o2[k][1])
        np.testing.assert_allclose(o2[k][0], o1[k][0], atol=1e-6)


def test_oracle_small_random_shape():
    """CUDA path vs the numpy oracle on a shape that has no golden file (T=10, 5x3, L=9)."""
    from vgqa_b200.engine import GroundingEngine
    seed, T, H, W, L = 5, 10, 5, 3, 9
    sd = O.synth_state_dict(seed)
    vis, vid, pos, text = O.synth_inputs(seed, T, H, W, L)
    ref = O.hot_path_forward(sd, vis, vid, pos, text, return_debug=True)
    eng = GroundingEngine(sd, max_clips=1, max_frames=16, max_hw=16, max_text=16)
    w1 = np.zeros(T, np.float32); w1[ref["debug"]["choose_pass1"]] = 1
    w2 = np.zeros(T, np.float32); w2[ref["debug"]["choose_pass2"]] = 1
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    o = eng.forward(t(vis[None]), t(vid[None]), t(text[None, :, 0]), t(pos[:1]), force_choose1=t(w1[None]),
                    force_choose2=t(w2[None]), want=["pred_boxes", "pred_sted", "logits_f_m", "logits_r_m", "encoded_feature"])
    torch.cuda.synchronize()
    assert float(np.abs(o["pred_boxes"][0].cpu().numpy() - ref["pred_boxes"]).max()) <= TOL
    assert float(np.abs(o["pred_sted"][0].cpu().numpy() - ref["pred_sted"][0]).max()) <= TOL
    assert float(np.abs(o["logits_f_m"][0].cpu().numpy() - ref["logits_f_m"]).max()) <= TOL
    enc = o["encoded_feature"].cpu().numpy()              # [T, S, 256] frame-major
    ref_enc = ref["debug"]["encoded_feature"].transpose(1, 0, 2)
    assert float(np.abs(enc - ref_enc).max()) <= 6e-2    # bf16 activations of O(1..4) magnitude
    eng.close()


def test_error_conventions():
    from vgqa_b200.engine import GroundingEngine
    sd = O.synth_state_dict(0)
    eng = GroundingEngine(sd, max_clips=1, max_frames=8, max_hw=9, max_text=8)
    z = lambda *s: torch.zeros(*s, device="cuda")
    with pytest.raises(RuntimeError, match="capacity"):
        eng.forward(z(1, 16, 256, 3, 3), z(1, 16, 256, 3, 3), z(1, 4, 256), z(1, 256, 3, 3))
    with pytest.raises(RuntimeError, match="bad shape"):
        eng.forward(z(1, 1, 256, 3, 3), z(1, 1, 256, 3, 3), z(1, 4, 256), z(1, 256, 3, 3))
    eng.close()
    sd.pop("bbox_embed.layers.2.weight")
    with pytest.raises(RuntimeError, match="missing weight"):
        GroundingEngine(sd, max_clips=1, max_frames=8, max_hw=9, max_text=8)
