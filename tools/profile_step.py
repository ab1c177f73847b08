"""One warm-up + N eager (no CUDA graph) steps of the bench workload, for ncu.  usage: profile_step.py [clips] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vgqa_oracle as O  # synthetic weights / inputs only
from vgqa_b200.engine import GroundingEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
T, H, W, L = 64, 7, 7, 20
RAW = os.environ.get("VGQA_PROFILE_RAW") == "1"   # enter the path at the extractor outputs: raw maps + RoBERTa token ids
CH, TOWER = (2048, 768, 768), (12, 50265)
sd = O.synth_state_dict(0, front_end_ch=CH, text_tower=TOWER) if RAW else O.synth_state_dict(0)
eng = GroundingEngine(sd, max_clips=B, max_frames=T, max_hw=H * W, max_text=L, use_cuda_graph=False)
base = [O.synth_inputs(i, T, H, W, L) for i in range(4)]
vis = torch.from_numpy(np.stack([base[i % 4][0] for i in range(B)])).cuda()
vid = torch.from_numpy(np.stack([base[i % 4][1] for i in range(B)])).cuda()
text = torch.from_numpy(np.stack([base[i % 4][3][:, 0, :] for i in range(B)])).cuda()
pos = torch.from_numpy(base[0][2][:1].copy()).cuda()
sizes = torch.tensor([[360.0, 640.0]] * B).cuda()
outs = eng.alloc_outputs(B, T, H, W, L, ["pred_boxes", "pred_sted", "boxes_px", "sted_idx"])
if RAW:
    g = torch.Generator(device="cuda").manual_seed(1)
    vis = torch.randn(B, T, CH[0], H, W, device="cuda", generator=g).relu_()
    vid = torch.randn(B, T, CH[1], H, W, device="cuda", generator=g)
    ids = torch.from_numpy(O.synth_text_ids(0, B, L, TOWER[1])[0]).cuda()
for _ in range(1 + steps):
    if RAW:
        eng.forward(vis, vid, None, pos, ori_sizes_hw=sizes, outs=outs, raw=True, text_ids=ids)
    else:
        eng.forward(vis, vid, text, pos, ori_sizes_hw=sizes, outs=outs)
    torch.cuda.synchronize()
print("launches per step:", eng.last_launch_count)
