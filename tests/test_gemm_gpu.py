"""GPU: the tcgen05 GEMM (+fused epilogues) through the C-ABI against a torch fp32 reference.
Tolerance: bf16 operands, fp32 accumulate → |err| <= 2e-2 * (1 + |ref|) on O(1) data."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(A, W, C, bias=None, bias_period=1, act=0, mul=None, res=None, ln=None, eps=1e-5):
    from vgqa_b200 import _lib
    L = _lib.lib()
    M, K = A.shape
    N = W.shape[0]
    st = torch.cuda.current_stream().cuda_stream
    bias_ld = 0 if bias is None else (bias.stride(0) if bias.dim() == 2 else N)
    _lib.check(L.vgqa_gemm_bf16(
        _lib.ptr(A), A.stride(0), _lib.ptr(W), W.stride(0), M, N, K, _lib.ptr(C), C.stride(0),
        1 if C.dtype == torch.float32 else 0, _lib.ptr(bias), bias_period, bias_ld,
        act, _lib.ptr(mul), 0 if mul is None else mul.stride(0), _lib.ptr(res), 0 if res is None else res.stride(0),
        _lib.ptr(ln[0]) if ln else None, _lib.ptr(ln[1]) if ln else None, eps, st))
    torch.cuda.synchronize()


def _ref(A, W, bias=None, bias_period=1, act=0, mul=None, res=None, ln=None, eps=1e-5):
    y = A.float() @ W.float().t()
    if bias is not None:
        if bias.dim() == 2:
            idx = torch.arange(A.shape[0], device=A.device) % bias_period
            y = y + bias[idx]
        else:
            y = y + bias
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    if mul is not None:
        y = y * mul.float()
    if res is not None:
        y = y + res.float()
    if ln:
        y = torch.nn.functional.layer_norm(y, (y.shape[1],), ln[0], ln[1], eps)
    return y


def _close(out, ref, tol=2e-2):
    err = (out.float() - ref).abs()
    lim = tol * (1 + ref.abs())
    assert bool((err <= lim).all()), f"max err {err.max().item():.4g}, worst ratio {(err / lim).max().item():.3g}"


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (300, 768, 256), (7552, 2048, 256),
                                   (1000, 256, 2048), (64, 64, 128), (33, 128, 512), (4096, 1536, 256)])
def test_gemm_plain(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    C = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    _gemm(A, W, C, bias=b)
    _close(C, _ref(A, W, b))


def test_gemm_narrow_tiles_many_per_cta_on_fresh_memory():
    """N = 320 → 64-column tiles with ONE bf16 store unit per tile and 13 tiles per CTA: the epilogue's two staging slabs must
    alternate per store, not per unit index (a slab was once rewritten while its previous TMA store was still reading it; the
    race showed only on slow first-touch stores — hence the freshly allocated outputs).  Shape of the first Video-Swin stage."""
    M, N, K = 50176, 320, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    ref = _ref(A, W, b)
    first = None
    for _ in range(6):
        torch.cuda.empty_cache()                     # the next output comes from a new cudaMalloc
        C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        _gemm(A, W, C, bias=b)
        _close(C, ref)
        if first is None:
            first = C.clone()
        assert torch.equal(C, first), "the GEMM must be bit-reproducible"
        del C


def test_gemm_fp32_out_table_relu():
    g = torch.Generator(device="cuda").manual_seed(3)
    M, N, K, period = 1180, 768, 256, 118
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / 16).bfloat16()
    tab = torch.randn(period, N, device="cuda", generator=g)
    C = torch.zeros(M, N, device="cuda")
    _gemm(A, W, C, bias=tab, bias_period=period, act=1)
    _close(C, _ref(A, W, tab, period, act=1), tol=1e-2)


@pytest.mark.parametrize("K", [256, 2048])
@pytest.mark.parametrize("eps", [1e-5, 1e-12])
def test_gemm_res_layernorm(K, eps):
    g = torch.Generator(device="cuda").manual_seed(K)
    M, N = 777, 256
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    res = (torch.randn(M, N, device="cuda", generator=g) + 3.0).bfloat16()   # non-zero mean stresses the variance
    lw = 1 + 0.1 * torch.randn(N, device="cuda", generator=g)
    lb = 0.1 * torch.randn(N, device="cuda", generator=g)
    C = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    _gemm(A, W, C, bias=b, res=res, ln=(lw, lb), eps=eps)
    _close(C, _ref(A, W, b, res=res, ln=(lw, lb), eps=eps))


def test_gemm_gelu_mul_strided():
    g = torch.Generator(device="cuda").manual_seed(11)
    M, N, K = 640, 256, 512
    Abig = torch.randn(M, 768, device="cuda", generator=g).bfloat16()
    A = Abig[:, 256:]                       # strided A view (lda = 768)
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    mul = torch.randn(M, 512, device="cuda", generator=g).bfloat16()[:, :256]
    Cbig = torch.zeros(M, 768, device="cuda", dtype=torch.bfloat16)
    C = Cbig[:, 512:]
    _gemm(A, W, C, bias=b, act=2, mul=mul)
    _close(C, _ref(A, W, b, act=2, mul=mul))
    assert float(Cbig[:, :512].abs().max()) == 0.0


@pytest.mark.parametrize("M,K,res,c32,c2,act", [(1000, 256, True, True, True, 0),
                                                (80000, 256, True, True, True, 0),      # several tiles per SM: the persistent loop wraps
                                                (80000 + 77, 256, True, True, False, 0),  # ragged last tile
                                                (3000, 2048, True, True, False, 0),     # few rows, long K: the CTA-pair (column-split) form
                                                (76000, 512, False, False, False, 1),   # no residual, ReLU, bf16 output only
                                                (77000, 64, True, False, True, 0),
                                                (64, 2048, True, True, True, 0),        # one row tile: the four-CTA cluster form (batch-1 decoders)
                                                (100, 256, True, True, False, 1),       # the same with a short K and ReLU
                                                (128, 512, False, False, True, 0)])
def test_gemm_residual_layernorm_all_outputs(M, K, res, c32, c2, act):
    """vgqa_gemm_ln (the `norm(x + sublayer(x))` launches of the forward) vs torch fp32: bf16 / fp32 / bf16(x+pos) outputs."""
    import ctypes
    from vgqa_b200 import _lib
    L = _lib.lib()
    L.vgqa_gemm_ln.restype = ctypes.c_int
    v = ctypes.c_void_p
    L.vgqa_gemm_ln.argtypes = [v, ctypes.c_int, v, ctypes.c_int, ctypes.c_int, ctypes.c_int, v, ctypes.c_int, v, v, v, ctypes.c_float,
                               v, v, v, v, ctypes.c_int, v]
    g = torch.Generator(device="cuda").manual_seed(M + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(256, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(256, device="cuda", generator=g) * 0.1
    r = torch.randn(M, 256, device="cuda", generator=g) if res else None
    lw = 1 + 0.1 * torch.randn(256, device="cuda", generator=g)
    lb = 0.1 * torch.randn(256, device="cuda", generator=g)
    period = 118
    add2 = torch.randn(period, 256, device="cuda", generator=g).bfloat16()
    C = torch.full((M + 4, 256), 9.0, device="cuda", dtype=torch.bfloat16)      # canary rows behind the end
    C32 = torch.full((M + 4, 256), 9.0, device="cuda") if c32 else None
    C2 = torch.full((M + 4, 256), 9.0, device="cuda", dtype=torch.bfloat16) if c2 else None
    _lib.check(L.vgqa_gemm_ln(_lib.ptr(A), K, _lib.ptr(W), K, M, K, _lib.ptr(b), act, _lib.ptr(r), _lib.ptr(lw), _lib.ptr(lb), 1e-5,
                              _lib.ptr(C), _lib.ptr(C32), _lib.ptr(C2), _lib.ptr(add2), period, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    y = A.float() @ W.float().t() + b
    if act == 1:
        y = torch.relu(y)
    if res:
        y = y + r
    ref = torch.nn.functional.layer_norm(y, (256,), lw, lb, 1e-5)
    _close(C[:M], ref)
    assert float((C[M:].float() - 9.0).abs().max()) == 0.0
    if c32:
        assert float((C32[:M] - ref).abs().max()) <= 2e-2 and float((C32[M:] - 9.0).abs().max()) == 0.0
    if c2:
        idx = torch.arange(M, device="cuda") % period
        _close(C2[:M], ref + add2.float()[idx])
        assert float((C2[M:].float() - 9.0).abs().max()) == 0.0
