import sys, os
R0 = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R0)
import numpy as np, torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from ptt_b200 import ops, train_ops as T
DEV="cuda:0"
rs = np.random.RandomState(0)
def rel(a,b,name):
    a=a.double(); b=b.double(); print("  %-10s rel err %.2e" % (name, float((a-b).abs().max()/b.abs().max())))
for (groups, ns, C0, C1) in ((1024,32,128,128),(1024,1,128,128),(500,32,128,128),(1024,32,128,256),(1024,16,128,128)):
    R = groups*ns
    print("groups",groups,"ns",ns,"R",R, C0, C1)
    x = torch.from_numpy(rs.standard_normal((R,C0)).astype(np.float32)).to(DEV)
    W0 = torch.from_numpy((rs.standard_normal((C0,C0))/np.sqrt(C0)).astype(np.float32)).to(DEV)
    W1 = torch.from_numpy((rs.standard_normal((C1,C0))/np.sqrt(C0)).astype(np.float32)).to(DEV)
    g0 = torch.from_numpy(rs.uniform(.5,1.5,C0).astype(np.float32)).to(DEV); b0 = torch.from_numpy(rs.normal(0,.3,C0).astype(np.float32)).to(DEV)
    g1 = torch.from_numpy(rs.uniform(.5,1.5,C1).astype(np.float32)).to(DEV); b1 = torch.from_numpy(rs.normal(0,.3,C1).astype(np.float32)).to(DEV)
    dout = torch.from_numpy(rs.standard_normal((groups,C1)).astype(np.float32)).to(DEV)
    # torch
    xt = x.clone().requires_grad_(True)
    y0 = xt @ W0.t(); y0.retain_grad()
    z0 = torch.relu(F.batch_norm(y0.t().reshape(1,C0,R), None, None, g0, b0, training=True)[0].t()); z0.retain_grad()
    y1 = z0 @ W1.t(); y1.retain_grad()
    z1 = torch.relu(F.batch_norm(y1.t().reshape(1,C1,R), None, None, g1, b1, training=True)[0].t())
    out = z1.reshape(groups, ns, C1).max(1)[0]
    out.backward(dout)
    # native
    l0 = ops.PackedLinear(W0); ny0 = l0(x)
    ka0,kb0,m0,r0 = T.bn_train_finalize(T.col_stats(ny0,C0), R, g0, b0, 1e-5, 0.1, None, None)
    l1 = ops.PackedLinear(W1); ny1 = l1(ny0, in_affine=(ka0,kb0))
    ka1,kb1,m1,r1 = T.bn_train_finalize(T.col_stats(ny1,C1), R, g1, b1, 1e-5, 0.1, None, None)
    nout, arg = T.bn_relu_maxpool(ny1, groups, ns, C1, ka1, kb1)
    rel(nout, out, "out")
    dy1, s = T.bn_relu_bwd(dout, arg, ns, ny1, C1, ka1, kb1, m1, r1, g1)
    rel(dy1, y1.grad, "dy1")
    dz0 = ops.PackedLinear(W1.t().contiguous())(dy1)
    rel(dz0, z0.grad, "dz0")
    dz0b = ops.PackedLinear(W1.t().contiguous())(y1.grad.contiguous())
    rel(dz0b, z0.grad, "dz0(torch dy1)")
    dy0, s0 = T.bn_relu_bwd(dz0, None, 1, ny0, C0, ka0, kb0, m0, r0, g0)
    rel(dy0, y0.grad, "dy0")
    dy0b, s0 = T.bn_relu_bwd(z0.grad.contiguous(), None, 1, ny0, C0, ka0, kb0, m0, r0, g0)
    rel(dy0b, y0.grad, "dy0(torch dz0)")
