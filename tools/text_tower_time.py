"""Device time of the RoBERTa-base text tower + resizer (vgqa_text_tower) for Q queries of L tokens.  usage: text_tower_time.py [Q L]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vgqa_oracle as O  # synthetic weights / ids only
from vgqa_b200.engine import GroundingEngine

Q, L = (int(a) for a in sys.argv[1:3]) if len(sys.argv) >= 3 else (64, 20)
layers, vocab = 12, 50265
sd = O.synth_state_dict(0, front_end_ch=(2048, 768, 768), text_tower=(layers, vocab))
eng = GroundingEngine(sd, max_clips=Q, max_frames=4, max_hw=4, max_text=L)
ids, _ = O.synth_text_ids(0, Q, L, vocab)
tids = torch.from_numpy(ids).cuda()
for _ in range(3):
    eng.text_tower(tids)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 10
for _ in range(n):
    eng.text_tower(tids)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
R = Q * L
flops = layers * 2.0 * R * (4 * 768 * 768 + 2 * 768 * 3072) + 2.0 * R * 768 * 256
print(f"text tower (RoBERTa-base, {layers} layers, vocab {vocab}) + resizer: {Q} queries x {L} tokens: {ms * 1e3:.0f} us "
      f"({flops / ms / 1e9:.1f} TFLOP/s, {ms * 1e3 / Q:.1f} us per query; eager launches)")
