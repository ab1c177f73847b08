"""ResNet101 extractor (csrc/resnet.cu) on `clips` x 64 frames @224: warm-up + timed runs (CUDA events), and — with --torch — the same
network in PyTorch / cuDNN (fp32 NCHW, and bf16 autocast channels_last) beside it, with the deviation of each from the fp32 run.
For ncu: `python tools/resnet_once.py 8 --once`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vgqa_b200 import synth as O
from vgqa_b200.engine import GroundingEngine

clips = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
once = "--once" in sys.argv
rsd = O.synth_resnet101(0)
sd = O.synth_state_dict(0)
sd.update(rsd)
eng = GroundingEngine(sd, max_clips=1, max_frames=8, max_hw=49, max_text=8)
x = torch.randn(clips * 64, 3, 224, 224, device="cuda")


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        y = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, y


if once:
    for _ in range(2):
        eng.resnet_backbone(x)
        torch.cuda.synchronize()
    print("launches:", eng.last_launch_count)
    sys.exit(0)
ms, y = timed(lambda: eng.resnet_backbone(x), 5)
flops = clips * 64 * 2 * 7.8e9
print(f"library: {ms:.2f} ms per {clips * 64} frames ({eng.last_launch_count} launches), {flops / ms / 1e9:.0f} TFLOP/s of the network's 7.8 GMAC/frame")
if "--torch" in sys.argv:
    import torchvision
    from torchvision.ops.misc import FrozenBatchNorm2d      # same arithmetic as the reference's class (eps 1e-5)
    net = torchvision.models.resnet101(weights=None, norm_layer=FrozenBatchNorm2d)
    net.load_state_dict({k[len("vis_encoder.0.body."):]: torch.from_numpy(v) for k, v in rsd.items()}, strict=False)
    body = torch.nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool, net.layer1, net.layer2, net.layer3, net.layer4).cuda().eval()
    with torch.no_grad():
        torch.backends.cudnn.benchmark = True
        ms32, y32 = timed(lambda: body(x), 3)
        bcl = body.to(memory_format=torch.channels_last)
        xcl = x.to(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms16, y16 = timed(lambda: bcl(xcl), 3)
    ref = y32.permute(0, 2, 3, 1).float()
    e_lib = (y.float() - ref).abs()
    e_ac = (y16.permute(0, 2, 3, 1).float() - ref).abs()
    print(f"PyTorch cuDNN fp32 NCHW: {ms32:.2f} ms   bf16 autocast channels_last: {ms16:.2f} ms")
    print(f"|layer4| mean {ref.abs().mean().item():.3f}; deviation from the fp32 run — library: mean {e_lib.mean().item():.4f} max {e_lib.max().item():.3f};"
          f" torch bf16 autocast: mean {e_ac.mean().item():.4f} max {e_ac.max().item():.3f}")
