import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, time
from ptt_b200 import train_ops as T, ops
R, M, N = 393216, 256, 128
dy = torch.randn(R, M, device="cuda"); x = torch.randn(R, N, device="cuda")
ka = torch.rand(N, device="cuda") + .5; kb = torch.randn(N, device="cuda")
for _ in range(3): T.linear_wgrad(dy, x, M, N, (ka, kb))
torch.cuda.synchronize()
for (r, m, n) in ((393216,256,128),(393216,128,128),(98304,512,512),(786432,64,64),(786432,128,64)):
    dy = torch.randn(r, m, device="cuda"); x = torch.randn(r, n, device="cuda")
    T.linear_wgrad(dy, x, m, n); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): T.linear_wgrad(dy, x, m, n)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print("wgrad R=%d M=%d N=%d: %.1f us  %.1f TFLOP/s(alg)  min-traffic %.0f GB/s" % (r, m, n, ms*1e3, 2.0*r*m*n/ms/1e9, (r*(m+n)*4)/ms/1e6))
    y = torch.randn(r, m, device="cuda")
    T.col_stats(y, m); torch.cuda.synchronize()
    e0.record()
    for _ in range(5): T.col_stats(y, m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print("  col_stats R=%d C=%d: %.1f us %.0f GB/s" % (r, m, ms*1e3, r*m*4/ms/1e6))
    w = torch.randn(n, m, device="cuda")/16
    lin = ops.PackedLinear(w)
    lin(y); torch.cuda.synchronize()
    e0.record()
    for _ in range(5): lin(y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print("  tc_gemm R=%d K=%d N=%d: %.1f us %.1f TFLOP/s(alg) traffic %.0f GB/s" % (r, m, n, ms*1e3, 2.0*r*m*n/ms/1e9, r*(m+n)*4/ms/1e6))
    # BatchNorm + ReLU backward over the same rows (dense form): reads dz, y; writes dy
    dz = torch.randn(r, m, device="cuda")
    ka_, kb_ = torch.rand(m, device="cuda") + .5, torch.randn(m, device="cuda")
    mean, rstd, gamma = torch.randn(m, device="cuda"), torch.rand(m, device="cuda") + .5, torch.rand(m, device="cuda") + .5
    T.bn_relu_bwd(dz, None, 1, y, m, ka_, kb_, mean, rstd, gamma); torch.cuda.synchronize()
    e0.record()
    for _ in range(5): T.bn_relu_bwd(dz, None, 1, y, m, ka_, kb_, mean, rstd, gamma)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print("  bn_relu_bwd (reduce + apply) R=%d C=%d: %.1f us  %.0f GB/s (5 tensor passes)" % (r, m, ms*1e3, 5.0*r*m*4/ms/1e6))
