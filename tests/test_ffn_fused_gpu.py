"""GPU: the fused FFN block (CTA-pair tcgen05 kernel, ffn_fused.cu) through the C-ABI against a torch fp32 reference
of  LayerNorm(x32 + W2 relu(W1 x + b1) + b2)  (modal_encoder.py:175-177).
Tolerance: bf16 operands / bf16 hidden activation, fp32 accumulation → |err| <= 2e-2 * (1 + |ref|) on O(1) data."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, F, res=True, c32=True, c2=True, period=118, parts=0, seed=0):
    from vgqa_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(seed + M + F)
    x32 = torch.randn(M, 256, device="cuda", generator=g)
    x = x32.bfloat16()
    W1 = (torch.randn(F, 256, device="cuda", generator=g) / 16).bfloat16()
    b1 = torch.randn(F, device="cuda", generator=g) * 0.1
    W2 = (torch.randn(256, F, device="cuda", generator=g) / F ** 0.5).bfloat16()
    b2 = torch.randn(256, device="cuda", generator=g) * 0.1
    lw = 1 + 0.1 * torch.randn(256, device="cuda", generator=g)
    lb = 0.1 * torch.randn(256, device="cuda", generator=g)
    pos = torch.randn(period, 256, device="cuda", generator=g).bfloat16()
    C = torch.zeros(M, 256, device="cuda", dtype=torch.bfloat16)
    C32 = torch.zeros(M, 256, device="cuda") if c32 else None
    C2 = torch.zeros(M, 256, device="cuda", dtype=torch.bfloat16) if c2 else None
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.vgqa_ffn_fused(_lib.ptr(x), _lib.ptr(W1), _lib.ptr(b1), _lib.ptr(W2), _lib.ptr(b2), M, F,
                                _lib.ptr(x32) if res else None, _lib.ptr(lw), _lib.ptr(lb), 1e-5, _lib.ptr(C),
                                _lib.ptr(C32), _lib.ptr(C2), _lib.ptr(pos) if c2 else None, period, parts, st))
    torch.cuda.synchronize()
    h = torch.relu(x.float() @ W1.float().t() + b1).bfloat16().float()
    y = h @ W2.float().t() + b2
    if res:
        y = y + x32
    y = torch.nn.functional.layer_norm(y, (256,), lw, lb, 1e-5)

    def close(out, ref, what):
        err = (out.float() - ref).abs()
        lim = 2e-2 * (1 + ref.abs())
        assert bool((err <= lim).all()), f"{what}: max err {err.max().item():.4g}, worst ratio {(err / lim).max().item():.3g}"

    close(C, y, "C")
    if c32:
        close(C32, y, "C32")
        assert (C32 - y).abs().max().item() < 1.5e-2
    if c2:
        idx = torch.arange(M, device="cuda") % period
        close(C2, y + pos[idx].float(), "C2")


@pytest.mark.parametrize("M,F", [(256, 256), (256, 2048), (128, 2048), (300, 2048), (7552, 2048), (3776, 1024), (1, 256)])
def test_ffn_fused(M, F):
    _run(M, F)


def test_ffn_fused_many_tiles_per_pair():
    _run(64 * 118 * 40, 2048)      # 1180 pair-tiles over 74 pairs: the persistent loop wraps 16 times


def test_ffn_fused_optional_outputs():
    _run(1000, 2048, res=False, c32=False, c2=False)
    _run(1000, 2048, res=True, c32=True, c2=True, period=1000)
    _run(1000, 512, res=True, c32=False, c2=True, period=7)
