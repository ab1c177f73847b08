"""In-kernel timeline of the fused FFN kernel (build with VGQA_EXTRA_NVCC_FLAGS=-DVGQA_FFN_PROFILE):
per chunk slot of CTA pair 0: MMA-thread wait times on h_ready / the weight ring, and the epilogue's turn-around."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vgqa_b200 import _lib

L = _lib.lib()
M, F = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 64 * 118, 2048
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 4
x32 = torch.randn(M, 256, device="cuda"); x = x32.bfloat16()
W1 = (torch.randn(F, 256, device="cuda") / 16).bfloat16(); W2 = (torch.randn(256, F, device="cuda") / F ** 0.5).bfloat16()
b1 = torch.zeros(F, device="cuda"); b2 = torch.zeros(256, device="cuda"); lw = torch.ones(256, device="cuda")
pos = torch.randn(118, 256, device="cuda").bfloat16()
C = torch.empty(M, 256, device="cuda", dtype=torch.bfloat16); C2 = torch.empty_like(C); C32 = torch.empty(M, 256, device="cuda")
for _ in range(3):
    _lib.check(L.vgqa_ffn_fused(_lib.ptr(x), _lib.ptr(W1), _lib.ptr(b1), _lib.ptr(W2), _lib.ptr(b2), M, F, _lib.ptr(x32), _lib.ptr(lw),
                                _lib.ptr(b2), 1e-5, _lib.ptr(C), _lib.ptr(C32), _lib.ptr(C2), _lib.ptr(pos), 118, parts,
                                torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32768)()
L.vgqa_ffn_prof_read(buf)
nchunk = F // 128
for blk in (0, 1):
    rows = [[buf[blk * 16384 + g * 16 + k] for k in range(16)] for g in range(1024)]
    lead = [[buf[g * 16 + k] for k in range(16)] for g in range(1024)]
    print(f"---- CTA {blk} (times relative to the leader's slot start)")
    print("slot  period wait_hready g2f0  g2f1  g1f0  g1f1 | hidden pass: hacc_full  arrive | LN pass: out_full  out_empty  end")
    prev = None
    for g in range(14, 52):
        r = rows[g]; s0 = lead[g][0]
        per = s0 - prev if prev is not None else 0
        prev = s0
        rel = lambda v: (v - s0) if v else -1
        ln = f"{rel(r[12]):7d} {rel(r[13]):7d} res {rel(r[9]):7d} mom {rel(r[10]):7d} slab0 {rel(r[11]):7d} end {rel(r[14]):7d}" if g % nchunk == 0 else ""
        print(f"{g:4d} {per:6d} {lead[g][1] - s0:6d} {rel(lead[g][4]):6d} {rel(lead[g][5]):6d} {rel(lead[g][2]):6d} {rel(lead[g][3]):6d}  | "
              f"{rel(r[6]):7d} {rel(r[7]):7d} | {ln}")
