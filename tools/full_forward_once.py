"""One full forward from pixels (32 frames @224) on the library: timing of its parts (CUDA events)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vgqa_b200 import synth as O
from vgqa_b200.engine import GroundingEngine

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sd = O.synth_state_dict(0, front_end_ch=(2048, 768, 768), text_tower=(12, 50265))
sd.update(O.synth_resnet101(0)); sd.update(O.synth_swin_backbone(0))
ids = torch.tensor([[0, 5910, 36777, 43933, 5772, 33081, 34763, 2]], dtype=torch.int32, device="cuda")
eng = GroundingEngine(sd, max_clips=1, max_frames=T, max_hw=49, max_text=8, use_cuda_graph=True)
x = torch.randn(T, 3, 224, 224, device="cuda")
outs = eng.alloc_outputs(1, T, 7, 7, 8, ["pred_boxes", "pred_sted"])


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


vis, vid = eng.extract_features(x, 1)
print(f"frames {T}: resnet {timed(lambda: eng.resnet_backbone(x)):.2f} ms ({eng.last_launch_count} launches)  swin {timed(lambda: eng.swin_backbone(x, 1)):.2f} ms"
      f"  both (two streams) {timed(lambda: eng.extract_features(x, 1)):.2f} ms"
      f"  forward from maps + ids {timed(lambda: eng.forward(vis, vid, None, None, outs=outs, raw=True, text_ids=ids)):.2f} ms")
