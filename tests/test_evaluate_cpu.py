"""CPU: the evaluation harness (vgqa_b200/evaluate.py) against golden I/O of the reference's own
`VidSTGEvaluator` (tests/golden/make_golden_eval.py), the even/odd merge of `do_eval`, and the cross-rank merge on gloo (world 2)."""
import json
import os

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN
from vgqa_b200 import evaluate as E


def load_golden():
    g = json.load(open(os.path.join(GOLDEN, "eval_golden.json")))
    gt = [{**d, "bboxs": {int(k): v for k, v in d["bboxs"].items()}} for d in g["gt"]]
    preds = {int(v): {int(f): b for f, b in p.items()} for v, p in g["predictions"].items()}
    vpreds = {int(k): v for k, v in g["video_predictions"].items()}
    kf = {int(k): v for k, v in g["kf"].items()}
    return g, gt, preds, vpreds, kf


def test_metrics_match_reference_evaluator():
    g, gt, preds, vpreds, kf = load_golden()
    ev = E.VidSTGEvaluator(gt, [0.3, 0.5])
    ev.update(preds); ev.update_kf_pr(kf); ev.video_update(vpreds)
    out = ev.summarize()
    assert set(out) == set(g["summary"])
    for k, v in g["summary"].items():
        assert out[k] == pytest.approx(v, abs=1e-12), k
    for vid, m in g["per_video"].items():
        r = ev.results[int(vid)]
        assert r["qtype"] == m["qtype"]
        for n in ("tiou", "viou", "gt_viou"):
            assert float(r[n]) == pytest.approx(m[n], abs=1e-12)


def test_iou_and_edge_cases():
    a = np.array([[0.0, 0.0, 10.0, 10.0]]); b = np.array([[5.0, 5.0, 15.0, 15.0], [20.0, 20.0, 30.0, 30.0]])
    np.testing.assert_allclose(E.np_box_iou(a, b), [[25.0 / 175.0, 0.0]])
    ev = E.VidSTGiouEvaluator([{"item_id": 1, "gt_temp_bound": [4, 8], "bboxs": {4: [0, 0, 1, 1]}}])
    m, _, _ = ev.evaluate({}, {1: {"sted": [8, 12], "qtype": "none"}}, {}, {1: (0.5, 0.25)})   # touching segments, no boxes
    assert m[1]["tiou"] == 0 and m[1]["viou"] == 0 and m[1]["gt_viou"] == 0 and m[1]["kf_pr"] == (0.5, 0.25)
    assert E.precision_recall([], [1, 2]) == (0, 0.0) and E.precision_recall([1, 2, 3], [2, 3, 4, 5]) == (2 / 3, 0.5)


def test_merge_even_odd_follows_do_eval():
    p1 = ({7: {0: [[0.0, 0.0, 10.0, 10.0]], 4: [[4.0, 4.0, 14.0, 14.0]]}}, {7: {0: [0.2], 4: [0.6]}},
          {7: {"sted": [0, 5], "qtype": "inter"}}, {7: (1.0, 0.5)})
    p2 = ({7: {2: [[2.0, 2.0, 12.0, 12.0]], 6: [[6.0, 6.0, 16.0, 16.0]]}}, {7: {2: [0.4], 6: [0.8]}},
          {7: {"sted": [2, 7], "qtype": "inter"}}, {7: (0.5, 0.0)})
    bbox, att, temp, kf = E.merge_even_odd(p1, p2)
    assert sorted(bbox[7]) == list(range(7)) and bbox[7][1] == [[1.0, 1.0, 11.0, 11.0]] and bbox[7][5] == [[5.0, 5.0, 15.0, 15.0]]
    assert [att[7][f][0] for f in range(7)] == [0.2, 0.2, 0.4, 0.4, 0.6, 0.6, 0.8]     # left value up to interval // 2
    assert temp[7] == {"sted": [0, 7], "qtype": "inter"} and kf[7] == [0.75, 0.25]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, gt, preds, vpreds, kf = load_golden()
    ids = sorted(preds)
    lo, hi = E.partition_clips(len(ids), world, rank)
    ev = E.VidSTGEvaluator(gt, [0.3, 0.5])
    for v in ids[lo:hi]:                               # every rank contributes only its own shard
        ev.update({v: preds[v]}); ev.update_kf_pr({v: kf[v]}); ev.video_update({v: vpreds[v]})
    ev.synchronize_between_processes()
    out = ev.summarize(main_process=(rank == 0))
    q.put((rank, out, len(ev.predictions)))
    dist.destroy_process_group()


def test_cross_rank_merge_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = dict((r, (o, n)) for r, o, n in (q.get(timeout=120) for _ in ps))
    [p.join(60) for p in ps]
    g = load_golden()[0]
    assert res[1][0] is None and res[0][1] == res[1][1] == len(g["predictions"])
    for k, v in g["summary"].items():
        assert res[0][0][k] == pytest.approx(v, abs=1e-12), k


class _FakeEngine:
    """Deterministic stand-in for GroundingEngine.forward (CPU): lets `do_eval`'s batching / partition / gather logic run under
    gloo without a GPU.  Outputs depend only on each clip's own inputs, like the real engine's."""
    device = "cpu"

    def forward(self, vis, vid, text, pos, ori_sizes_hw=None, want=None, raw=False):
        import torch
        B, T = vis.shape[:2]
        m = vis.reshape(B, T, -1).mean(-1) + vid.reshape(B, T, -1).mean(-1)                    # [B, T]
        box = torch.stack([100 + 50 * m, 80 + 40 * m, 300 + 50 * m, 260 + 40 * m], -1)
        idx = torch.stack([m.argmin(1).clamp(max=T - 2), torch.full((B,), T - 1)], 1).to(torch.int32)
        return {"boxes_px": box, "att_sequences": torch.sigmoid(m), "sted_idx": idx, "choose2": (m > 0).float()}


def _fake_items(n=7, T2=8):
    rng = np.random.Generator(np.random.PCG64(3))
    items, gt = [], []
    for i in range(n):
        fids = list(range(2 * i, 2 * i + T2))
        items.append({"item_id": 500 + i, "vis": rng.standard_normal((T2, 4, 2, 2)).astype(np.float32),
                      "vid": rng.standard_normal((T2, 4, 2, 2)).astype(np.float32), "text": np.zeros((3, 4), np.float32),
                      "pos": np.zeros((1, 256, 2, 2), np.float32), "frame_ids": fids, "ori_size": (360, 640),
                      "qtype": ["declar", "inter"][i % 2], "actioness": (np.arange(T2) % 3 == 0).astype(np.float32)})
        gt.append({"item_id": 500 + i, "gt_temp_bound": [fids[1], fids[-2]], "bboxs": {f: [110.0, 90.0, 310.0, 270.0] for f in fids[1:-2]}})
    return items, gt


def _eval_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    items, gt = _fake_items()
    ev = E.VidSTGEvaluator(gt, [0.3, 0.5])
    out = E.do_eval(_FakeEngine(), items, ev, clips_per_call=2, rank=rank, world=world)
    q.put((rank, out, ev.video_predictions))
    dist.destroy_process_group()


def test_do_eval_two_ranks_equals_one_rank_gloo():
    items, gt = _fake_items()
    ref_ev = E.VidSTGEvaluator(gt, [0.3, 0.5], distributed=False)
    ref = E.do_eval(_FakeEngine(), items, ref_ev, clips_per_call=3)
    assert len(ref_ev.video_predictions) == len(items) and set(k.split("_")[0] for k in ref) == {"declar", "inter"}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    ps = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = dict((r, (o, vp)) for r, o, vp in (q.get(timeout=120) for _ in ps))
    [p.join(60) for p in ps]
    assert res[1][0] is None                                    # only rank 0 summarizes
    assert res[0][1] == res[1][1] == ref_ev.video_predictions    # every rank holds the merged predictions
    for k, v in ref.items():
        assert res[0][0][k] == pytest.approx(v, abs=1e-12), k
