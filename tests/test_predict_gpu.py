"""GPU: the persistent predict() pipeline (vgqa_b200/predict.py): even/odd passes as one two-clip batch, device PostProcess,
merge — against the same pipeline done the reference's way (two separate single-clip forwards + PostProcess + merge), and
against the oracle's merge on the reference golden of the even pass."""
import numpy as np
import pytest
import torch

from conftest import golden_path
from oracle import vgqa_oracle as O
from vgqa_b200.engine import GroundingEngine
from vgqa_b200.predict import GroundingPredictor
from vgqa_b200 import postprocess as PP

pytestmark = pytest.mark.gpu


def _inputs(seed, n, H, W, L):
    vis, vid, pos, text = O.synth_inputs(seed, n, H, W, L)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t(vis), t(vid), t(pos), t(text[:, 0])


def _reference_way(sd, vis, vid, pos, text, fids, ori, fps, T, H, W, L):
    """grounding.py:180-244 with the hot path swapped in: one forward per parity, PostProcess module, merge."""
    eng = GroundingEngine(sd, max_clips=1, max_frames=T, max_hw=H * W, max_text=L)
    post = PP.PostProcess()
    passes = []
    for par in (0, 1):
        o = eng.forward(vis[par::2][None].contiguous(), vid[par::2][None].contiguous(), text[None].contiguous(), pos[:1].contiguous())
        outputs = {"pred_sted": o["pred_sted"], "pred_boxes": o["pred_boxes"][0], "att_sequences": o["att_sequences"], "pr": (0, 0)}
        sizes = torch.tensor([list(ori)] * T, device="cuda")
        pf = fids[par::2]
        boxes, att, steds, _ = post(outputs, sizes, [pf], [T])
        b, a = boxes.cpu().tolist(), att.reshape(-1).cpu().tolist()
        passes.append(({0: {pf[j]: [b[j]] for j in range(T)}}, {0: {pf[j]: [a[j]] for j in range(T)}},
                       {0: {"sted": steds[0], "qtype": "declar"}}, {}))
    eng.close()
    return PP.merge_predictions(passes[0], passes[1], fps)


def test_predictor_matches_two_separate_forwards():
    seed, T, H, W, L = 1, 32, 7, 7, 20
    sd = O.synth_state_dict(seed)
    vis, vid, pos, text = _inputs(seed, 2 * T, H, W, L)
    fids = list(range(10, 10 + 2 * T * 3, 3))            # 2T sampled frame ids, stride 3
    ori, fps = (360, 640), 25.0
    pred = GroundingPredictor(sd, sample_num=T, max_hw=H * W, max_text=L, max_queries=2)
    got = pred.predict(vis, vid, text, pos, fids, ori, fps)
    ref = _reference_way(sd, vis, vid, pos, text, fids, ori, fps, T, H, W, L)
    assert set(got) == {"temporal", "tube"} and got["temporal"]["score"] == 1.0
    assert got["temporal"] == ref["temporal"]
    assert [t["frame"] for t in got["tube"]] == [t["frame"] for t in ref["tube"]] == list(range(fids[0], fids[-1] + 1))
    gb = np.asarray([t["bbox"] for t in got["tube"]]); rb = np.asarray([t["bbox"] for t in ref["tube"]])
    np.testing.assert_allclose(gb, rb, atol=0.5)          # pixels; the batched forward runs the same kernels on 2 clips
    np.testing.assert_allclose([t["score"] for t in got["tube"]], [t["score"] for t in ref["tube"]], atol=2e-2)
    # two queries in one call: same answer for the repeated item, independent of its slot in the batch
    both = pred.predict_many([{"vis": vis, "vid": vid, "text": text, "pos": pos, "frame_ids": fids, "ori_size": ori, "fps": fps}] * 2)
    assert both[0]["temporal"] == both[1]["temporal"] == got["temporal"]
    np.testing.assert_allclose(np.asarray([t["bbox"] for t in both[1]["tube"]]), gb, atol=0.5)
    pred.close()


def test_predictor_even_pass_matches_reference_golden():
    g = np.load(golden_path("cfg1_T32_7x7_L20_s1"))
    T, H, W, L, seed = (int(g[k]) for k in ("T", "H", "W", "L", "seed"))
    sd = O.synth_state_dict(seed)
    vis, vid, pos, text = _inputs(seed, T, H, W, L)
    # interleave the golden clip with itself: even and odd passes both see the golden frames
    il = lambda x: torch.stack([x, x], 1).reshape((2 * x.shape[0],) + tuple(x.shape[1:]))
    fe = g["frame_ids"].tolist()
    fids = sorted(fe + [f + 1 for f in fe])
    pred = GroundingPredictor(sd, sample_num=T, max_hw=H * W, max_text=L)
    got = pred.predict(il(vis), il(vid), text, pos, fids, tuple(int(v) for v in g["ori_size"]), fps=1.0)
    s, e = g["post_sted"][0].tolist()
    assert got["temporal"]["start"] == float(s) and got["temporal"]["end"] == float(e + 1)   # odd pass = even pass shifted by one frame
    by_frame = {t["frame"]: t["bbox"] for t in got["tube"]}
    np.testing.assert_allclose(np.asarray([by_frame[f] for f in fe]), g["post_boxes"], atol=2e-2 * 640)
    pred.close()


def test_predictor_argument_checks():
    sd = O.synth_state_dict(0)
    pred = GroundingPredictor(sd, sample_num=8, max_hw=16, max_text=8)
    vis, vid, pos, text = _inputs(0, 7, 4, 4, 6)
    with pytest.raises(ValueError):
        pred.predict(vis, vid, text, pos, list(range(7)), (10, 10))          # odd number of sampled frames
    with pytest.raises(ValueError):
        pred.predict_many([{}] * 2)                                           # more queries than max_queries
    pred.close()


def test_fresh_tensors_do_not_recapture_graphs():
    """CUDA graphs are keyed on the shape of a call, never on the caller's pointers: ten predict() calls on freshly allocated
    tensors capture each phase exactly once (two graphs), and `pos` may be left to the library."""
    seed, T, H, W, L = 2, 16, 7, 7, 12
    sd = O.synth_state_dict(seed)
    pred = GroundingPredictor(sd, sample_num=T, max_hw=H * W, max_text=L, use_cuda_graph=True)
    fids = list(range(0, 2 * T))
    first = None
    keep = []
    for i in range(10):
        vis, vid, pos, text = _inputs(seed, 2 * T, H, W, L)          # new device tensors on every call
        keep.append((vis, vid, text))                                # keep them alive: the allocator cannot hand the addresses out again
        got = pred.predict(vis, vid, text, None if i % 2 else pos, fids, (360, 640), 25.0)
        if first is None:
            first = got
            captured = pred.engine.graph_capture_count
            assert captured == 2, captured
        else:
            assert got["temporal"] == first["temporal"]
            np.testing.assert_allclose(np.asarray([t["bbox"] for t in got["tube"]]), np.asarray([t["bbox"] for t in first["tube"]]),
                                       atol=0.5)     # pixels; the library-generated pos differs from the given one by fp32 rounding
    assert pred.engine.graph_capture_count == 2, "a new tensor address must not trigger a re-capture"
    pred.close()
