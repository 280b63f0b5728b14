import sys, os
R0 = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R0)
import numpy as np, torch, torch.nn.functional as F
from ptt_b200 import ops, train_ops as T
DEV="cuda:0"
rs = np.random.RandomState(0)
for (groups, ns, C) in ((1024,1,128),(1024,32,256),(1024,32,128),(1024,2,128),(64,1,128),(2048,1,128),(1024,32,512),(512,32,256)):
    R = groups*ns
    y = torch.from_numpy(rs.standard_normal((R,C)).astype(np.float32)).to(DEV)
    g = torch.from_numpy(rs.uniform(.5,1.5,C).astype(np.float32)).to(DEV); b = torch.from_numpy(rs.normal(0,.3,C).astype(np.float32)).to(DEV)
    dout = torch.from_numpy(rs.standard_normal((groups,C)).astype(np.float32)).to(DEV)
    yt = y.clone().requires_grad_(True)
    z = torch.relu(F.batch_norm(yt.t().reshape(1,C,R), None, None, g, b, training=True)[0].t())
    out = z.reshape(groups, ns, C).max(1)[0]
    out.backward(dout)
    ka,kb,m,r = T.bn_train_finalize(T.col_stats(y,C), R, g, b, 1e-5, 0.1, None, None)
    nout, arg = T.bn_relu_maxpool(y, groups, ns, C, ka, kb)
    dy, s = T.bn_relu_bwd(dout, arg, ns, y, C, ka, kb, m, r, g)
    torch.cuda.synchronize()
    d = (dy - yt.grad).abs()
    bad = (d > 1e-4).nonzero()
    # reference sums
    mask = (z > 0).float()
    mm = torch.zeros_like(y); idx = z.reshape(groups,ns,C).max(1)[1]
    mm.view(groups,ns,C).scatter_(1, idx[:,None,:], dout[:,None,:]); mm = mm*mask
    yh = (y - m)*r
    s1 = mm.double().sum(0); s2 = (mm*yh).double().sum(0)
    print(groups, ns, C, "max err %.2e nbad %d" % (float(d.max()), len(bad)), "s1 err %.2e s2 err %.2e" % (float((s[0]-s1).abs().max()), float((s[1]-s2).abs().max())),
          "bad cols", sorted(set(bad[:,1].tolist()))[:12], "bad rows", sorted(set(bad[:,0].tolist()))[:8])
