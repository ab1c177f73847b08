"""A few launches of input_proj_kernel at the bench shape (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import bench_kernels as B

C = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B.bench_input_proj(C=C, second=(C != 2048), name="once")
torch.cuda.synchronize()
