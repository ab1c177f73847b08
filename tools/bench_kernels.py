"""Micro-benchmarks of single kernels through the C-ABI (CUDA events, L2-exceeding operands)."""
import math
import sys
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vgqa_b200 import _lib

L = _lib.lib()
st = lambda: torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def bench_attn(F=4096, S=118):
    qkv = torch.randn(F * S, 768, device="cuda").bfloat16()
    O = torch.empty(F * S, 256, device="cuda", dtype=torch.bfloat16)
    for use_tc in (1, 0):
        if use_tc and L.vgqa_enc_attn(_lib.ptr(qkv), _lib.ptr(O), F, S, None, 1 / math.sqrt(32), 1, st()) != 0:
            print(f"enc_attn tcgen05 F={F} S={S}: not supported"); continue
        t = timeit(lambda: _lib.check(L.vgqa_enc_attn(_lib.ptr(qkv), _lib.ptr(O), F, S, None, 1 / math.sqrt(32), use_tc, st())))
        flops = 4.0 * S * S * 32 * 8 * F
        byts = F * S * (768 + 256) * 2
        print(f"enc_attn {'tcgen05' if use_tc else 'mma.sync'} F={F} S={S}: {t:8.1f} us  {flops / t / 1e6:7.1f} TFLOP/s  {byts / t / 1e3:7.1f} GB/s")


def bench_gemm(M, N, K, ln=False, act=0, name=""):
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.zeros(N, device="cuda")
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    lw = torch.ones(N, device="cuda") if ln else None
    t = timeit(lambda: _lib.check(L.vgqa_gemm_bf16(_lib.ptr(A), K, _lib.ptr(W), K, M, N, K, _lib.ptr(C), N, 0, _lib.ptr(b), 1, N, act,
                                                   None, 0, None, 0, _lib.ptr(lw), _lib.ptr(b) if ln else None, 1e-5, st())))
    flops = 2.0 * M * N * K
    byts = (M * K + M * N + N * K) * 2
    print(f"gemm {name} M={M} N={N} K={K} ln={ln}: {t:8.1f} us  {flops / t / 1e6:7.1f} TFLOP/s  {byts / t / 1e3:7.1f} GB/s")


def bench_xattn(F=4096, S=118, tok0=0, Mk=69, use_pos=False, use_kpos=False, name=""):
    mem_all = torch.randn(F, S, 256, device="cuda").bfloat16()
    qt = (torch.randn(F, 8, 256, device="cuda") / 4).bfloat16()
    pos = torch.randn(Mk, 256, device="cuda").bfloat16() if use_pos else None
    q2 = torch.randn(F, 256, device="cuda").bfloat16() if use_kpos else None
    kpos = torch.randn(Mk, 1536, device="cuda").bfloat16() if use_kpos else None
    ctx = torch.zeros(F, 2048, device="cuda", dtype=torch.bfloat16)
    mem = mem_all[:, tok0:tok0 + Mk]
    t = timeit(lambda: _lib.check(L.vgqa_xattn1(_lib.ptr(qt), _lib.ptr(mem), S, F, Mk, _lib.ptr(pos), 0, _lib.ptr(q2), _lib.ptr(kpos), 1536,
                                                0, None, 0, 0.125, _lib.ptr(ctx), None, st())))
    byts = F * (Mk * 512 + 4096 + 4096)
    print(f"xattn1 {name} F={F} Mk={Mk}: {t:8.1f} us  {byts / t / 1e3:7.1f} GB/s")


def bench_xattn_stream(F=4096, S=118, tok0=0, Mk=69, bias=True, name=""):
    mem_all = torch.randn(F, S, 256, device="cuda").bfloat16()
    qt = (torch.randn(F, 8, 256, device="cuda") / 4).bfloat16()
    ldsb = (Mk + 7) // 8 * 8
    sb = torch.randn(F, 8, ldsb, device="cuda") if bias else None
    ctx = torch.zeros(F, 2048, device="cuda", dtype=torch.bfloat16)
    mem = mem_all[:, tok0:tok0 + Mk]
    t = timeit(lambda: _lib.check(L.vgqa_xattn1_bias(_lib.ptr(qt), _lib.ptr(mem), S, F, Mk, _lib.ptr(sb), ldsb, None, 0, 0.125,
                                                     _lib.ptr(ctx), None, st())))
    byts = F * (Mk * 512 + 4096 + 4096 + (8 * Mk * 4 if bias else 0))
    print(f"xattn stream {name} F={F} Mk={Mk}: {t:8.1f} us  {byts / t / 1e3:7.1f} GB/s")


def bench_ffn(M, F=2048, parts=0):
    x32 = torch.randn(M, 256, device="cuda")
    x = x32.bfloat16()
    W1 = (torch.randn(F, 256, device="cuda") / 16).bfloat16()
    W2 = (torch.randn(256, F, device="cuda") / F ** 0.5).bfloat16()
    b1 = torch.zeros(F, device="cuda"); b2 = torch.zeros(256, device="cuda"); lw = torch.ones(256, device="cuda")
    pos = torch.randn(118, 256, device="cuda").bfloat16()
    C = torch.empty(M, 256, device="cuda", dtype=torch.bfloat16); C2 = torch.empty_like(C); C32 = torch.empty(M, 256, device="cuda")
    t = timeit(lambda: _lib.check(L.vgqa_ffn_fused(_lib.ptr(x), _lib.ptr(W1), _lib.ptr(b1), _lib.ptr(W2), _lib.ptr(b2), M, F, _lib.ptr(x32),
                                                   _lib.ptr(lw), _lib.ptr(b2), 1e-5, _lib.ptr(C), _lib.ptr(C32), _lib.ptr(C2), _lib.ptr(pos), 118, parts, st())))
    flops = 4.0 * M * F * 256
    byts = M * 3584.0
    print(f"ffn_fused M={M} F={F}: {t:8.1f} us  {flops / t / 1e6:7.1f} TFLOP/s  {byts / t / 1e3:7.1f} GB/s")


def bench_input_proj(F=4096, C=2048, H=7, W=7, Lt=20, second=False, name=""):
    import ctypes
    L.vgqa_input_proj.restype = ctypes.c_int
    L.vgqa_input_proj.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int] + [ctypes.c_void_p] * 3 + \
        [ctypes.c_int] * 4 + [ctypes.c_void_p]
    P = H * W
    S = 2 * P + Lt
    x = torch.randn(F, C, H, W, device="cuda")
    Wt = (torch.randn(256, C, device="cuda") / C ** 0.5).bfloat16()
    b = torch.zeros(256, device="cuda")
    pos = torch.randn(S, 256, device="cuda").bfloat16()
    X = torch.empty(F * S, 256, device="cuda", dtype=torch.bfloat16); XP = torch.empty_like(X); X32 = torch.empty(F * S, 256, device="cuda")
    t = timeit(lambda: _lib.check(L.vgqa_input_proj(_lib.ptr(x), C, _lib.ptr(Wt), _lib.ptr(b), _lib.ptr(pos), 1, _lib.ptr(X), _lib.ptr(X32),
                                                    _lib.ptr(XP), F, S, (P + Lt) if second else 0, P, st())))
    flops = 2.0 * F * P * 256 * C
    byts = F * P * (C * 4.0 + 2048)
    print(f"input_proj {name} F={F} C={C} P={P}: {t:8.1f} us  {flops / t / 1e6:7.1f} TFLOP/s  {byts / t / 1e3:7.1f} GB/s (algorithmic bytes)")
    L.vgqa_input_proj_nhwc.restype = ctypes.c_int
    L.vgqa_input_proj_nhwc.argtypes = L.vgqa_input_proj.argtypes
    x16 = x.permute(0, 2, 3, 1).contiguous().bfloat16()
    t = timeit(lambda: _lib.check(L.vgqa_input_proj_nhwc(_lib.ptr(x16), C, _lib.ptr(Wt), _lib.ptr(b), _lib.ptr(pos), 1, _lib.ptr(X),
                                                         _lib.ptr(X32), _lib.ptr(XP), F, S, (P + Lt) if second else 0, P, st())))
    byts = F * P * (C * 2.0 + 2048)
    print(f"input_proj {name} channels-last bf16 F={F} C={C} P={P}: {t:8.1f} us  {flops / t / 1e6:7.1f} TFLOP/s  {byts / t / 1e3:7.1f} GB/s (algorithmic bytes)")


def bench_gemm_ln(M, K=256, name=""):
    import ctypes
    v = ctypes.c_void_p
    L.vgqa_gemm_ln.restype = ctypes.c_int
    L.vgqa_gemm_ln.argtypes = [v, ctypes.c_int, v, ctypes.c_int, ctypes.c_int, ctypes.c_int, v, ctypes.c_int, v, v, v, ctypes.c_float,
                               v, v, v, v, ctypes.c_int, v]
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(256, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.zeros(256, device="cuda"); lw = torch.ones(256, device="cuda")
    r = torch.randn(M, 256, device="cuda")
    C = torch.empty(M, 256, device="cuda", dtype=torch.bfloat16); C32 = torch.empty(M, 256, device="cuda")
    t = timeit(lambda: _lib.check(L.vgqa_gemm_ln(_lib.ptr(A), K, _lib.ptr(W), K, M, K, _lib.ptr(b), 0, _lib.ptr(r), _lib.ptr(lw), _lib.ptr(b),
                                                 1e-5, _lib.ptr(C), _lib.ptr(C32), None, None, 1, st())))
    byts = M * (K * 2 + 1024 + 512 + 1024.0)
    print(f"gemm_ln {name} M={M} K={K} (residual in, bf16 + fp32 out): {t:8.1f} us  {2.0 * M * 256 * K / t / 1e6:7.1f} TFLOP/s  {byts / t / 1e3:7.1f} GB/s")


if __name__ == "__main__":
    which = sys.argv[1:] or ["attn", "gemm"]
    if "attn" in which:
        bench_attn()
        bench_attn(F=1024, S=118)
    if "attn_long" in which:
        bench_attn(F=1024, S=352)
        bench_attn(F=1024, S=412)
        bench_attn(F=1024, S=256)
    if "xattn_long" in which:
        bench_xattn_stream(F=1024, S=352, Mk=208, tok0=0, name="cfg-5 decoder (stream)")
        bench_xattn(F=1024, S=352, Mk=208, tok0=0, name="cfg-5 decoder (CTA per frame, no positional terms)")
        bench_xattn_stream(F=1024, S=352, Mk=144, tok0=208, bias=False, name="cfg-5 spatial (stream)")
        bench_xattn(F=1024, S=352, Mk=144, tok0=208, name="cfg-5 spatial (CTA per frame)")
    if "xattn" in which:
        bench_xattn_stream(Mk=49, tok0=69, bias=False, name="spatial")
        bench_xattn_stream(Mk=69, tok0=0, name="decoder")
        bench_xattn(Mk=49, tok0=69, name="spatial")
        bench_xattn(Mk=69, tok0=0, use_kpos=True, name="pos-decoder")
        bench_xattn(Mk=69, tok0=49, use_pos=True, name="time-decoder")
    if "input_proj" in which:
        bench_input_proj(C=2048, name="ResNet101 -> vis tokens")
        bench_input_proj(C=768, second=True, name="Video-Swin -> vid tokens")
        bench_input_proj(F=1024, C=2048, H=14, W=14, name="14x14")
    if "gemm_ln" in which:
        bench_gemm_ln(64 * 64 * 118, name="encoder out-proj + LN1")
    if "ffn" in which:
        bench_ffn(64 * 64 * 118)
        bench_ffn(16 * 64 * 118)
    if "gemm" in which:
        R = 64 * 64 * 118
        bench_gemm(R, 768, 256, name="qkv")
        bench_gemm(R, 2048, 256, act=1, name="ffn1")
        bench_gemm(R, 256, 256, ln=True, name="outproj+ln (no residual)")
        bench_gemm(R, 256, 2048, ln=True, name="ffn2+ln (no residual)")
        bench_gemm(4096, 2048, 256, name="dec qabs")
        bench_gemm(4096, 256, 2048, ln=True, name="dec vo+ln")
