"""A few launches of enc_attn_tc_long_kernel at S = 352 (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import bench_kernels as B

B.bench_attn(F=1024, S=352)
torch.cuda.synchronize()
