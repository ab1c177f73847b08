"""Golden I/O of the reference's VidSTG metrics (vgqa/data/metrics/vidstg_evaluator.py) on synthetic predictions.

    python tests/golden/make_golden_eval.py       # writes tests/golden/eval_golden.json   (build container only)

The REAL `VidSTGEvaluator` / `VidSTGiouEvaluator` classes are driven exactly as `do_eval` drives them
(update / update_att / update_kf_pr / video_update / summarize); their annotation cache is a temporary torch.save file."""
import json
import logging
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402


def synth(seed=11, n=7):
    rng = np.random.Generator(np.random.PCG64(seed))
    gt, preds, vpreds, kf = [], {}, {}, {}
    for i in range(n):
        item = 100 + i
        s = int(rng.integers(0, 20)); e = s + int(rng.integers(4, 30))
        boxes = {}
        for f in range(s, e):
            x, y = rng.uniform(0, 300, 2); w, h = rng.uniform(40, 200, 2)
            boxes[f] = [float(x), float(y), float(x + w), float(y + h)]
        gt.append({"item_id": item, "gt_temp_bound": [s, e], "bboxs": boxes, "description": f"query {i}"})
        ps = max(0, s + int(rng.integers(-6, 7))); pe = ps + int(rng.integers(2, 35))
        if i == 3:
            ps, pe = e + 2, e + 9                                   # disjoint prediction → tiou 0
        pb = {}
        for f in range(min(s, ps), max(e, pe)):
            if f in boxes and rng.uniform() < 0.9:
                g = np.asarray(boxes[f]); pb[f] = [(g + rng.normal(0, 25, 4)).tolist()]
            elif rng.uniform() < 0.5:
                pb[f] = [rng.uniform(0, 400, 4).tolist()]
        preds[item] = pb
        vpreds[item] = {"sted": [ps, pe], "qtype": ["declar", "inter", "none"][i % 3]}
        kf[item] = [float(rng.uniform()), float(rng.uniform())]
    return gt, preds, vpreds, kf


if __name__ == "__main__":
    load_reference()
    import importlib
    types = sys.modules
    import types as _t
    if "vgqa.data" not in sys.modules:
        from ref_loader import _namespace_pkg, REF_ROOT
        _namespace_pkg("vgqa.data", os.path.join(REF_ROOT, "vgqa", "data"))
        _namespace_pkg("vgqa.data.metrics", os.path.join(REF_ROOT, "vgqa", "data", "metrics"))
    ev = importlib.import_module("vgqa.data.metrics.vidstg_evaluator")
    gt, preds, vpreds, kf = synth()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "data_cache"))
        torch.save(gt, os.path.join(d, "data_cache", "vidstd-test-anno.cache"))
        E = ev.VidSTGEvaluator(logging.getLogger("golden"), d, "test", [0.3, 0.5])
    E.update(preds); E.update_kf_pr(kf); E.video_update(vpreds)
    out = E.summarize()
    rec = {"gt": [{**g, "bboxs": {str(k): v for k, v in g["bboxs"].items()}} for g in gt],
           "predictions": {str(v): {str(f): b for f, b in p.items()} for v, p in preds.items()},
           "video_predictions": {str(k): v for k, v in vpreds.items()}, "kf": {str(k): v for k, v in kf.items()},
           "summary": out,
           "per_video": {str(k): {n: (float(v[n]) if n not in ("qtype",) else v[n]) for n in ("tiou", "viou", "gt_viou", "qtype")}
                         for k, v in E.results.items()}}
    json.dump(rec, open(os.path.join(HERE, "eval_golden.json"), "w"))
    print(json.dumps(out, indent=1))
