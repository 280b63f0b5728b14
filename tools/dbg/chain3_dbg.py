import sys, os
R0 = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R0)
import numpy as np, torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from ptt_b200 import ops, train_ops as T
DEV="cuda:0"
rs = np.random.RandomState(0)
def rel(a,b,name):
    a=a.detach().double(); b=b.detach().double(); print("  %-22s rel err %.2e" % (name, float((a-b).abs().max()/b.abs().max())))
for (groups, ns, C0, C1) in ((1024,32,128,128),(1024,1,128,128),(1024,32,128,256)):
    R = groups*ns
    print("groups",groups,"ns",ns,"R",R, C0, C1)
    x = torch.from_numpy(rs.standard_normal((R,C0)).astype(np.float32)).to(DEV)
    W0 = torch.from_numpy((rs.standard_normal((C0,C0))/np.sqrt(C0)).astype(np.float32)).to(DEV)
    W1 = torch.from_numpy((rs.standard_normal((C1,C0))/np.sqrt(C0)).astype(np.float32)).to(DEV)
    g0 = torch.from_numpy(rs.uniform(.5,1.5,C0).astype(np.float32)).to(DEV); b0 = torch.from_numpy(rs.normal(0,.3,C0).astype(np.float32)).to(DEV)
    g1 = torch.from_numpy(rs.uniform(.5,1.5,C1).astype(np.float32)).to(DEV); b1 = torch.from_numpy(rs.normal(0,.3,C1).astype(np.float32)).to(DEV)
    dout = torch.from_numpy(rs.standard_normal((groups,C1)).astype(np.float32)).to(DEV)
    xt = x.clone().requires_grad_(True)
    y0 = xt @ W0.t(); y0.retain_grad()
    z0 = torch.relu(F.batch_norm(y0.t().reshape(1,C0,R), None, None, g0, b0, training=True)[0].t()); z0.retain_grad()
    y1 = z0 @ W1.t(); y1.retain_grad()
    z1 = torch.relu(F.batch_norm(y1.t().reshape(1,C1,R), None, None, g1, b1, training=True)[0].t())
    mx = z1.reshape(groups, ns, C1).max(1)
    mx[0].backward(dout)
    ny0 = ops.PackedLinear(W0)(x)
    ka0,kb0,m0,r0 = T.bn_train_finalize(T.col_stats(ny0,C0), R, g0, b0, 1e-5, 0.1, None, None)
    ny1 = ops.PackedLinear(W1)(ny0, in_affine=(ka0,kb0))
    rel(ny1, y1, "y1")
    print("   max |y1 diff| %.3e" % float((ny1 - y1).abs().max()))
    ka1,kb1,m1,r1 = T.bn_train_finalize(T.col_stats(ny1,C1), R, g1, b1, 1e-5, 0.1, None, None)
    rel(m1, y1.detach().mean(0), "mean1"); rel(r1, 1/torch.sqrt(y1.detach().var(0, unbiased=False)+1e-5), "rstd1")
    nout, arg = T.bn_relu_maxpool(ny1, groups, ns, C1, ka1, kb1)
    print("   argmax equal frac %.5f" % float((arg.long() == mx[1]).float().mean()))
    dy1, s = T.bn_relu_bwd(dout, arg, ns, ny1, C1, ka1, kb1, m1, r1, g1)
    rel(dy1, y1.grad, "dy1 (native y1, arg)")
    dy1b, s = T.bn_relu_bwd(dout, mx[1].int().contiguous(), ns, ny1, C1, ka1, kb1, m1, r1, g1)
    rel(dy1b, y1.grad, "dy1 (torch arg)")
    dy1c, s = T.bn_relu_bwd(dout, mx[1].int().contiguous(), ns, y1.detach().contiguous(), C1, ka1, kb1, m1, r1, g1)
    rel(dy1c, y1.grad, "dy1 (torch y1, arg)")
    zz = z1.detach().reshape(groups, ns, C1)
    ties = (zz == zz.max(1, keepdim=True)[0]).sum(1)
    print("   tied max frac %.4f ; max==0 frac %.4f" % (float((ties > 1).float().mean()), float((zz.max(1)[0]==0).float().mean())))
