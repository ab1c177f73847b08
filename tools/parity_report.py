"""Prints the max-abs error of every output of the CUDA path against every golden case (reference decisions
forced), plus free-running decision agreement.  Run on a B200: python tools/parity_report.py [> profiles/...]."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.chdir(ROOT)
from test_parity_gpu import CASES, EV_CASES, run_case  # noqa: E402

KEYS = ["pred_boxes", "pred_sted", "pred_actioness", "logits_f_m", "logits_f_a", "logits_r_a", "logits_r_m",
        "att_sequences", "aux_boxes", "aux_sted", "aux_actioness", "frames_cls", "actioness_pass1"]
print("| case | K1 / K2 of T | " + " | ".join(KEYS) + " | sted argmax | free-run choose1/choose2 |")
print("|---|---|" + "---|" * (len(KEYS) + 2))
for name in EV_CASES + CASES:
    g, o = run_case(name, force=name not in EV_CASES)       # the decisive fixtures are run with NOTHING forced
    single_pass = "iteration_rate" in g.files and int(g["iteration_rate"]) >= 0
    ref = {"pred_boxes": g["pred_boxes"], "pred_sted": g["pred_sted"][0], "pred_actioness": g["pred_actioness"][0, :, 0],
           "logits_f_m": g["logits_f_m"], "logits_f_a": g["logits_f_a"], "logits_r_a": g["logits_r_a"][0],
           "logits_r_m": g["logits_r_m"][0], "att_sequences": g["att_sequences"][0], "aux_boxes": g["aux_boxes"],
           "aux_sted": g["aux_sted"][:, 0], "aux_actioness": g["aux_actioness"][:, 0, :, 0], "frames_cls": g["frames_cls"],
           "actioness_pass1": g["actioness_pass1"]}
    errs = []
    for k in KEYS:
        if k == "actioness_pass1" and single_pass:
            errs.append(float("nan"))
            continue
        got = o[k]
        got = got[:, 0] if k.startswith("aux_") else (got if k == "frames_cls" else got[0])
        errs.append(float(np.abs(got.reshape(ref[k].shape) - ref[k]).max()))
    fid = g["frame_ids"]
    s, e = o["sted_idx"][0]
    sted_ok = [int(fid[s]), int(fid[e]) + 1] == g["post_sted"][0].tolist()
    _, of = run_case(name, force=False)
    T = int(g["T"])
    r1 = np.zeros(T); r1[g["choose_pass1"]] = 1
    r2 = np.zeros(T); r2[g["choose_pass2"]] = 1
    d1 = int((of["choose1"][0] != r1).sum()); d2 = int((of["choose2"][0] != r2).sum())
    print(f"| {name} | {len(g['choose_pass1'])} / {len(g['choose_pass2'])} of {T} | " + " | ".join(f"{x:.4f}" for x in errs) + f" | {'same' if sted_ok else 'DIFF'} (top2 gap {float(g['margin_sted_top2']):.3f}) | {d1}/{d2} frames differ (margins {float(g['margin_theta']):.4f}/{float(g['margin_act']):.4f}) |")
