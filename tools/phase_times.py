"""Wall-clock split of one step: encoder phase alone, full forward unpipelined (encoder + decoder back to back),
and the two-slot pipelined path bench.py measures.  usage: phase_times.py [clips [T H W L]]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vgqa_oracle as O  # synthetic weights / inputs only
from vgqa_b200.engine import GroundingEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T, H, W, L = (int(a) for a in sys.argv[2:6]) if len(sys.argv) >= 6 else (64, 7, 7, 20)
eng = GroundingEngine(O.synth_state_dict(0), max_clips=B, max_frames=T, max_hw=H * W, max_text=L, use_cuda_graph=True)
base = [O.synth_inputs(i, T, H, W, L) for i in range(4)]
vis = torch.from_numpy(np.stack([base[i % 4][0] for i in range(B)])).cuda()
vid = torch.from_numpy(np.stack([base[i % 4][1] for i in range(B)])).cuda()
text = torch.from_numpy(np.stack([base[i % 4][3][:, 0, :] for i in range(B)])).cuda()
pos = torch.from_numpy(base[0][2][:1].copy()).cuda()
sizes = torch.tensor([[360.0, 640.0]] * B).cuda()
want = ["pred_boxes", "pred_sted", "boxes_px", "sted_idx"]
outs = [eng.alloc_outputs(B, T, H, W, L, want) for _ in range(2)]


def timed(fn, n=6, drain=None):
    for _ in range(3):
        fn()
    if drain:
        drain()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    if drain:
        drain()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_enc = timed(lambda: eng.encode(vis, vid, text, pos))
t_full = timed(lambda: eng.forward(vis, vid, text, pos, ori_sizes_hw=sizes, outs=outs[0]))
i = [0]


def piped():
    s = i[0] & 1
    eng.forward_async(vis, vid, text, pos, ori_sizes_hw=sizes, outs=outs[s], slot=s)
    i[0] += 1


t_pipe = timed(piped, n=8, drain=lambda: (eng.wait(0), eng.wait(1)))
from vgqa_b200.engine import reference_flops
gf = reference_flops(T, H, W, L) / 1e9
print(f"T={T} {H}x{W} L={L}: {B / t_pipe * 1e3:.1f} clips/s = {B / t_pipe * gf:.0f} algorithmic TFLOP/s ({gf:.1f} GF per clip)")
print(f"clips={B}: encoder phase alone (eager, + fp32 copy of the encoded features) {t_enc:.2f} ms | "
      f"full forward unpipelined {t_full:.2f} ms | pipelined {t_pipe:.2f} ms per step")

if os.environ.get("VGQA_TIMELINE") == "1":
    import ctypes
    from vgqa_b200 import _lib
    L = _lib.lib()
    rows = []
    for k in range(8):
        piped()
        s_ = (i[0] - 1) & 1
        buf = (ctypes.c_float * 4)()
        # read the PREVIOUS call on the other slot (complete by now or soon), keeping two calls in flight
        if k >= 1:
            _lib.check(L.vgqa_debug_phase_times(eng._ctx, s_ ^ 1, buf))
            rows.append(list(buf))
    t0 = rows[0][0]
    print("step  enc_start  enc_end   dec_start  dec_end    (ms, relative)")
    for r in rows:
        print("      " + "  ".join(f"{x - t0:8.2f}" for x in r))
