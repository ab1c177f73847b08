"""A few launches of xattn_stream_kernel at the decoder shape (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import bench_kernels as B

B.bench_xattn_stream(Mk=69, tok0=0, name="decoder")
torch.cuda.synchronize()
