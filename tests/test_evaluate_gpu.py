"""GPU: `do_eval` (vgqa_b200/evaluate.py) over a small synthetic dataset — batched even/odd passes through the engine — against
one `GroundingPredictor.predict` call per item (tests/test_predict_gpu.py pins that path to the reference goldens)."""
import numpy as np
import pytest
import torch

from oracle import vgqa_oracle as O
from vgqa_b200 import evaluate as E
from vgqa_b200.engine import GroundingEngine
from vgqa_b200.predict import GroundingPredictor

pytestmark = pytest.mark.gpu


def make_items(n, T2, H, W, L):
    items, gt = [], []
    for i in range(n):
        vis, vid, pos, text = O.synth_inputs(50 + i, T2, H, W, L)
        fids = list(range(3 * i, 3 * i + T2))
        act = np.zeros(T2); act[T2 // 4: 3 * T2 // 4] = 1
        items.append({"item_id": 900 + i, "vis": vis, "vid": vid, "text": text[:, 0], "pos": pos[:1], "frame_ids": fids,
                      "ori_size": (360, 640), "qtype": ["declar", "inter"][i % 2], "actioness": act})
        gt.append({"item_id": 900 + i, "gt_temp_bound": [fids[T2 // 4], fids[3 * T2 // 4]],
                   "bboxs": {f: [100.0, 80.0, 300.0, 260.0] for f in fids[T2 // 4: 3 * T2 // 4]}})
    return items, gt


def test_do_eval_matches_per_item_predict():
    n, T2, H, W, L = 5, 16, 4, 4, 6
    sd = O.synth_state_dict(3)
    items, gt = make_items(n, T2, H, W, L)
    eng = GroundingEngine(sd, max_clips=4, max_frames=T2 // 2, max_hw=H * W, max_text=L)
    ev = E.VidSTGEvaluator(gt, [0.3, 0.5])
    out = E.do_eval(eng, items, ev, clips_per_call=2)        # batches of 2, 2, 1 items
    eng.close()
    assert set(ev.predictions) == {it["item_id"] for it in items}
    pred = GroundingPredictor(sd, sample_num=T2 // 2, max_hw=H * W, max_text=L, use_cuda_graph=False)
    for it in items:
        r = pred.predict(it["vis"], it["vid"], it["text"], it["pos"], it["frame_ids"], it["ori_size"], fps=1.0, qtype=it["qtype"])
        vid = it["item_id"]
        assert ev.video_predictions[vid]["sted"] == [int(r["temporal"]["start"]), int(r["temporal"]["end"])]
        assert ev.video_predictions[vid]["qtype"] == it["qtype"]
        assert sorted(ev.predictions[vid]) == [t["frame"] for t in r["tube"]]
        np.testing.assert_allclose(np.asarray([ev.predictions[vid][t["frame"]][0] for t in r["tube"]]),
                                   np.asarray([t["bbox"] for t in r["tube"]]), atol=1e-3)
        np.testing.assert_allclose([ev.att_predictions[vid][t["frame"]][0] for t in r["tube"]], [t["score"] for t in r["tube"]], atol=1e-6)
        p, q = ev.kf_pred[vid]
        assert 0 <= p <= 1 and 0 <= q <= 1
    pred.close()
    # summary: one entry per (qtype, metric); metrics recomputed from the gathered dicts by an independent evaluator
    ev2 = E.VidSTGEvaluator(gt, [0.3, 0.5])
    ev2.update(ev.predictions); ev2.update_kf_pr(ev.kf_pred); ev2.video_update(ev.video_predictions)
    assert out == ev2.summarize()
    assert {k.split("_")[0] for k in out} == {"declar", "inter"} and all(0 <= v <= 1 for v in out.values())
