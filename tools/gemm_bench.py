"""Tuning aid: time ptt_linear_fwd (tc_gemm) for the live row-block contraction shapes at B = 48."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptt_b200 import ops

shapes = [("G' sa2 search", 24576, 128, 128), ("G' sa3 search", 12288, 256, 128), ("cov_final", 6144, 256, 256),
          ("fc1", 6144, 256, 512), ("qkv", 6144, 512, 1536), ("fc2", 6144, 512, 256), ("box layer", 49152, 256, 256),
          ("box fc2", 3072, 512, 256), ("pairs 512", 98304, 512, 512)]
for name, R, K, N in shapes:
    x = torch.randn(R, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    lin = ops.PackedLinear(w, b)
    for _ in range(3):
        y = lin(x, relu=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = lin(x, relu=True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print("%-16s R=%6d K=%4d N=%5d  %8.1f us   %6.1f TFLOP/s (algorithmic)" % (name, R, K, N, us, 2.0 * R * K * N / us / 1e6))
