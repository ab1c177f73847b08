"""One warm-up + one run of the whole Video-Swin-T extractor (csrc/swin.cu) on 8 clips x 64 frames @224, for ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vgqa_b200 import synth as O
from vgqa_b200.engine import GroundingEngine

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sd = O.synth_state_dict(0)
sd.update(O.synth_swin_backbone(0))
eng = GroundingEngine(sd, max_clips=1, max_frames=8, max_hw=49, max_text=8)
x = torch.randn(clips * 64, 3, 224, 224, device="cuda")
for _ in range(2):
    eng.swin_backbone(x, clips)
    torch.cuda.synchronize()
print("launches:", eng.last_launch_count)
