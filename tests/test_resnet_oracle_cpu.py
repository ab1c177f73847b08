"""CPU: the numpy restatement of the ResNet101 extractor (oracle.resnet101_backbone: torchvision's Bottleneck network with the
reference's FrozenBatchNorm2d, backbone.py:13-57,104-113) against the goldens produced by that construction itself
(tests/golden/make_golden_resnet.py) — fp32 vs fp32."""
import numpy as np
import pytest

from conftest import golden_path
from make_golden_resnet import ROW_STEP, resnet_frames
from oracle import vgqa_oracle as O


def test_resnet101_oracle_matches_reference_golden():
    g = np.load(golden_path("resnet101_n3_64_s1"))
    n, R, seed = (int(g[k]) for k in ("n", "R", "seed"))
    outs = O.resnet101_backbone(O.synth_resnet101(seed), resnet_frames(seed, n, R))
    for l in range(4):
        scale = float(g[f"abs_mean{l}"])
        np.testing.assert_allclose(outs[l].reshape(-1, 256 << l)[::ROW_STEP], g[f"rows{l}"], atol=2e-4 * max(scale, 1.0), err_msg=f"layer {l + 1}")
    np.testing.assert_allclose(outs[3], g["y3"].astype(np.float32), atol=0.1, rtol=2e-3)     # the full map is stored in fp16


def test_frozen_bn_and_pool_primitives():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 4, 6, 6)).astype(np.float32)
    sd = {"b.weight": rng.uniform(0.5, 1.5, 4).astype(np.float32), "b.bias": rng.standard_normal(4).astype(np.float32),
          "b.running_mean": rng.standard_normal(4).astype(np.float32), "b.running_var": rng.uniform(0.5, 1.5, 4).astype(np.float32)}
    y = O.frozen_bn(x, sd, "b")
    ref = (x - sd["b.running_mean"][None, :, None, None]) / np.sqrt(sd["b.running_var"][None, :, None, None] + 1e-5) * \
        sd["b.weight"][None, :, None, None] + sd["b.bias"][None, :, None, None]
    np.testing.assert_allclose(y, ref, atol=1e-5)
    p = O.max_pool_3x3_s2(x)
    assert p.shape == (2, 4, 3, 3) and p[0, 0, 0, 0] == x[0, 0, :2, :2].max() and p[1, 3, 2, 2] == x[1, 3, 3:6, 3:6].max()
    w = rng.standard_normal((5, 4, 3, 3)).astype(np.float32)
    c = O.conv2d(x, w, stride=2, pad=1)
    assert c.shape == (2, 5, 3, 3)
    np.testing.assert_allclose(c[0, 2, 1, 1], (x[0, :, 1:4, 1:4] * w[2]).sum(), rtol=1e-4)


def test_resnet101_oracle_matches_the_reference_construction_live():
    """Where /root/reference (or its byte-compiled copy) and torchvision are present: the same construction run live."""
    from ref_loader import reference_modules_available
    if not reference_modules_available():
        pytest.skip("reference modules not available")
    torch = pytest.importorskip("torch")
    pytest.importorskip("torchvision")
    from make_golden_resnet import build_reference_body
    body = build_reference_body()
    seed, n, R = 3, 1, 64
    sd = O.synth_resnet101(seed, prefix="")
    body.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    x = resnet_frames(seed, n, R)
    with torch.no_grad():
        out = body(torch.from_numpy(x))
    mine = O.resnet101_backbone({("vis_encoder.0.body." + k): v for k, v in sd.items()}, x)
    for l in range(4):
        np.testing.assert_allclose(mine[l], out[str(l)].permute(0, 2, 3, 1).numpy(), atol=1e-3, err_msg=f"layer {l + 1}")
