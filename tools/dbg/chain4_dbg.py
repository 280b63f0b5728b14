import sys, os
R0 = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R0)
import numpy as np, torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from ptt_b200 import ops, synth, train_ops as T
DEV="cuda:0"
rs = np.random.RandomState(0)
def rel(a,b,name):
    a=a.detach().double(); b=b.detach().double(); print("  %-22s rel err %.2e" % (name, float((a-b).abs().max()/b.abs().max())))
B,N,C,M,ns,radius = 4,512,128,256,32,0.5
xyz = torch.from_numpy(synth.make_clouds(B, N, 600 + N, "dense", role="template")).to(DEV)
feats = torch.from_numpy(synth.features((B, C, N), seed=601 + N)).to(DEV)
inds, new_xyz = ops.furthest_point_sampling(xyz, M, return_new_xyz=True)
idx = ops.ball_query(new_xyz, xyz, radius, ns)
x0 = T.sa_group_rows(xyz, ops.cm_to_pm(feats), new_xyz, idx, C, radius, True)
R = x0.shape[0]; K0 = C+3
print("rows", R, "ld", x0.shape[1], "dup frac", float((idx[:,:,1:]==idx[:,:,:1]).float().mean()))
for mode in ("real", "random"):
    xx = x0 if mode == "real" else torch.from_numpy(rs.standard_normal(tuple(x0.shape)).astype(np.float32)).to(DEV)
    C0, C1 = 128, 128
    W0 = torch.from_numpy((rs.standard_normal((C0,K0))/np.sqrt(K0)).astype(np.float32)).to(DEV)
    W1 = torch.from_numpy((rs.standard_normal((C1,C0))/np.sqrt(C0)).astype(np.float32)).to(DEV)
    g0 = torch.from_numpy(rs.uniform(.5,1.5,C0).astype(np.float32)).to(DEV); b0 = torch.from_numpy(rs.normal(0,.3,C0).astype(np.float32)).to(DEV)
    g1 = torch.from_numpy(rs.uniform(.5,1.5,C1).astype(np.float32)).to(DEV); b1 = torch.from_numpy(rs.normal(0,.3,C1).astype(np.float32)).to(DEV)
    dout = torch.from_numpy(rs.standard_normal((R//ns,C1)).astype(np.float32)).to(DEV)
    xt = xx[:, :K0].clone().requires_grad_(True)
    W0t = W0.clone().requires_grad_(True)
    y0 = xt @ W0t.t(); y0.retain_grad()
    z0 = torch.relu(F.batch_norm(y0.t().reshape(1,C0,R), None, None, g0, b0, training=True)[0].t()); z0.retain_grad()
    y1 = z0 @ W1.t(); y1.retain_grad()
    z1 = torch.relu(F.batch_norm(y1.t().reshape(1,C1,R), None, None, g1, b1, training=True)[0].t())
    mx = z1.reshape(R//ns, ns, C1).max(1)
    mx[0].backward(dout)
    print("mode", mode)
    ny0 = ops.PackedLinear(W0)(xx)
    ka0,kb0,m0,r0 = T.bn_train_finalize(T.col_stats(ny0,C0), R, g0, b0, 1e-5, 0.1, None, None)
    ny1 = ops.PackedLinear(W1)(ny0, in_affine=(ka0,kb0))
    ka1,kb1,m1,r1 = T.bn_train_finalize(T.col_stats(ny1,C1), R, g1, b1, 1e-5, 0.1, None, None)
    nout, arg = T.bn_relu_maxpool(ny1, R//ns, ns, C1, ka1, kb1)
    rel(nout, mx[0], "out"); print("   argmax equal frac %.5f" % float((arg.long() == mx[1]).float().mean()))
    dy1, s = T.bn_relu_bwd(dout, arg, ns, ny1, C1, ka1, kb1, m1, r1, g1)
    rel(dy1, y1.grad, "dy1")
    dy1t, s = T.bn_relu_bwd(dout, mx[1].int().contiguous(), ns, ny1, C1, ka1, kb1, m1, r1, g1)
    rel(dy1t, y1.grad, "dy1 (torch arg)")
    dz0 = ops.PackedLinear(W1.t().contiguous())(dy1)
    rel(dz0, z0.grad, "dz0")
    dy0, s0 = T.bn_relu_bwd(dz0, None, 1, ny0, C0, ka0, kb0, m0, r0, g0)
    rel(dy0, y0.grad, "dy0")
    dW0 = T.linear_wgrad(dy0, xx, C0, K0)
    rel(dW0, W0t.grad, "dW0")
    dx0 = ops.PackedLinear(W0.t().contiguous())(dy0, ld_out=132)
    rel(dx0[:, :K0], xt.grad, "dx0")
