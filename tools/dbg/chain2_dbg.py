import sys, os
R0 = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R0)
import numpy as np, torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from ptt_b200 import ops, train_ops as T
DEV="cuda:0"
rs = np.random.RandomState(0)
def rel(a,b,name):
    a=a.detach().double(); b=b.detach().double(); print("  %-18s rel err %.2e" % (name, float((a-b).abs().max()/b.abs().max())))
for (groups, ns, C0, C1) in ((1024,1,128,128),(1024,32,128,256)):
    R = groups*ns
    print("groups",groups,"ns",ns,"R",R, C0, C1)
    x = torch.from_numpy(rs.standard_normal((R,C0)).astype(np.float32)).to(DEV)
    W1 = torch.from_numpy((rs.standard_normal((C1,C0))/np.sqrt(C0)).astype(np.float32)).to(DEV)
    g1 = torch.from_numpy(rs.uniform(.5,1.5,C1).astype(np.float32)).to(DEV); b1 = torch.from_numpy(rs.normal(0,.3,C1).astype(np.float32)).to(DEV)
    dout = torch.from_numpy(rs.standard_normal((groups,C1)).astype(np.float32)).to(DEV)
    for mode in ("gemm_y", "torch_y"):
        if mode == "gemm_y":
            ny1 = ops.PackedLinear(W1)(x)
        else:
            ny1 = (x @ W1.t()).contiguous()
        yt = ny1.clone().requires_grad_(True)
        z1 = torch.relu(F.batch_norm(yt.t().reshape(1,C1,R), None, None, g1, b1, training=True)[0].t())
        out = z1.reshape(groups, ns, C1).max(1)
        out[0].backward(dout)
        ka1,kb1,m1,r1 = T.bn_train_finalize(T.col_stats(ny1,C1), R, g1, b1, 1e-5, 0.1, None, None)
        nout, arg = T.bn_relu_maxpool(ny1, groups, ns, C1, ka1, kb1)
        print(" mode", mode, "argmax equal frac %.4f" % float((arg.long() == out[1]).float().mean()))
        rel(nout, out[0], "out")
        dyA, sA = T.bn_relu_bwd(dout, arg, ns, ny1, C1, ka1, kb1, m1, r1, g1)
        dyB, sB = T.bn_relu_bwd(dout, arg, ns, ny1, C1, ka1, kb1, m1, r1, g1)
        rel(dyA, dyB, "determinism")
        rel(dyA, yt.grad, "dy1")
        # with torch's argmax
        dyC, sC = T.bn_relu_bwd(dout, out[1].int().contiguous(), ns, ny1, C1, ka1, kb1, m1, r1, g1)
        rel(dyC, yt.grad, "dy1 (torch argmax)")
        # number of exact ties at max
        zz = z1.detach().reshape(groups, ns, C1)
        ties = (zz == zz.max(1, keepdim=True)[0]).sum(1)
        print("  groups with tied max: %.4f ; max==0 frac %.4f" % (float((ties > 1).float().mean()), float((zz.max(1)[0]==0).float().mean())))
