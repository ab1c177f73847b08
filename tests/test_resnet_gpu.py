"""GPU: the ResNet101 extractor on this library's kernels (csrc/resnet.cu: every convolution a tcgen05 GEMM — the 3x3 ones as nine
row-shifted blocks of the padded channels-last operand — with FrozenBN folded and bias / residual / ReLU in the epilogue) against the
goldens of the reference's own construction (vgqa/core/vision/backbone.py:13-57,104-113; tests/golden/make_golden_resnet.py)."""
import numpy as np
import pytest
import torch

from conftest import golden_path
from make_golden_resnet import ROW_STEP, resnet_frames
from oracle import vgqa_oracle as O

pytestmark = pytest.mark.gpu

# bf16 operands and bf16 activations between the convolutions (fp32 accumulation).  Yardstick: the deviation of torch's OWN
# bf16-autocast run of the same network from its fp32 run, stored per layer in the fixture — the CUDA path may not be worse than
# MEAN_X / MAX_X times that.
MEAN_X, MAX_X = 1.5, 2.0


def _engine(seed, frames):
    from vgqa_b200.engine import GroundingEngine
    sd = O.synth_state_dict(0)
    sd.update(O.synth_resnet101(seed))
    return GroundingEngine(sd, max_clips=1, max_frames=8, max_hw=49, max_text=8)


@pytest.mark.parametrize("name", ["resnet101_n3_64_s1", "resnet101_n2_224_s0"])
def test_resnet101_matches_reference_golden(name):
    g = np.load(golden_path(name))
    n, R, seed = (int(g[k]) for k in ("n", "R", "seed"))
    x = torch.from_numpy(resnet_frames(seed, n, R)).cuda()
    eng = _engine(seed, x)
    out, layers = eng.resnet_backbone(x, want_layers=True)
    torch.cuda.synchronize()
    for l in range(4):
        got = layers[l].cpu().numpy().reshape(-1, 256 << l)[::ROW_STEP]
        err = np.abs(got - g[f"rows{l}"])
        am, ax = float(g[f"autocast_err_mean{l}"]), float(g[f"autocast_err_max{l}"])
        assert float(err.mean()) <= MEAN_X * am, (l, float(err.mean()), am)
        assert float(err.max()) <= MAX_X * ax, (l, float(err.max()), ax)
    y3 = g["y3"].astype(np.float32)
    err = np.abs(out.float().cpu().numpy() - y3)
    assert float(err.mean()) <= MEAN_X * float(g["autocast_err_mean3"]) + 2e-3 * float(g["abs_mean3"]), float(err.mean())   # + bf16 output, fp16 fixture
    assert out.shape == (n, R // 32, R // 32, 2048) and out.dtype == torch.bfloat16
    assert eng.last_launch_count >= 3 + 33 * 3 + 4
    eng.close()


def test_resnet101_chunks_and_frames_are_independent():
    """More than one pass (a pass holds as many frames as the GPU has SMs): every frame's map equals the map of the same picture
    computed in a small batch, bit for bit (a row's accumulation order does not depend on its tile or on its neighbours)."""
    seed, R = 1, 64
    base = torch.from_numpy(resnet_frames(seed, 3, R)).cuda()
    eng = _engine(seed, base)
    small = eng.resnet_backbone(base)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    n = 2 * sms + 5                                              # three passes, the last one short
    big = eng.resnet_backbone(base.repeat((n + 2) // 3, 1, 1, 1)[:n].contiguous())
    torch.cuda.synchronize()
    for i in (0, 1, 2, sms - 1, sms, sms + 1, 2 * sms - 1, 2 * sms, n - 1):
        assert torch.equal(big[i], small[i % 3]), i
    eng.close()


def test_resnet101_feeds_the_raw_forward():
    """The layer4 map (channels-last bf16) IS the `vis_raw` (raw_layout = 1) input of the forward."""
    from vgqa_b200.engine import GroundingEngine
    seed, T, H, W, L = 2, 8, 7, 7, 6
    ch = (2048, 128, 64)
    sd = O.synth_state_dict(seed, front_end_ch=ch)
    sd.update(O.synth_resnet101(seed))
    eng = GroundingEngine(sd, max_clips=1, max_frames=T, max_hw=H * W, max_text=L)
    frames = torch.from_numpy(resnet_frames(seed, T, 224)).cuda()
    vis_map = eng.resnet_backbone(frames).reshape(1, T, H, W, 2048)
    g = torch.Generator(device="cuda").manual_seed(3)
    vid_map = torch.randn(1, T, H, W, ch[1], device="cuda", generator=g).to(torch.bfloat16)
    text = torch.randn(1, L, ch[2], device="cuda", generator=g)
    o = eng.forward(vis_map, vid_map, text, None, raw=True, want=["pred_boxes", "pred_sted"])
    torch.cuda.synchronize()
    assert torch.isfinite(o["pred_boxes"]).all() and torch.isfinite(o["pred_sted"]).all()
    eng.close()


@pytest.mark.parametrize("n,R", [(2, 32), (1, 96), (1, 160)])
def test_resnet101_other_resolutions_match_the_oracle(n, R):
    """Frame sides that are not 224 (odd map sides after the strides: 96 → 24 / 12 / 6 / 3, 160 → 40 / 20 / 10 / 5; 32 → a 1x1 layer4
    map) against the numpy oracle on the same seeded frames: relative deviation of every layer output of the order of bf16 (1 %)."""
    seed = 5
    x = resnet_frames(seed, n, R)
    ref = O.resnet101_backbone(O.synth_resnet101(seed), x)
    eng = _engine(seed, None)
    out, layers = eng.resnet_backbone(torch.from_numpy(x).cuda(), want_layers=True)
    torch.cuda.synchronize()
    for l in range(4):
        got = layers[l].cpu().numpy()
        assert got.shape == ref[l].shape
        rel = float(np.abs(got - ref[l]).mean() / np.abs(ref[l]).mean())
        assert rel <= 1.5e-2, (l, rel)
    eng.close()
