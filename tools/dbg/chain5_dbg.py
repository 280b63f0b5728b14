import sys, os
R0 = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R0); sys.path.insert(0, os.path.join(R0, "tests"))
import numpy as np, torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from ptt_b200 import ops, synth, modules, train_ops as T
from test_oracle_golden import sa_state_dict
DEV="cuda:0"
def rel(a,b,name):
    a=a.detach().double(); b=b.detach().double(); print("  %-26s rel err %.2e" % (name, float((a-b).abs().max()/b.abs().max())))
B,N,C,M,ns,radius = 4,512,128,256,32,0.5
mlp=[128,128,128]
sd = {k: torch.from_numpy(v) for k, v in synth.fill_state_dict(sa_state_dict(mlp), seed=500 + N).items()}
xyz = torch.from_numpy(synth.make_clouds(B, N, 600 + N, "dense", role="template")).to(DEV)
feats = torch.from_numpy(synth.features((B, C, N), seed=601 + N)).to(DEV)
rec = {}
orig_bwd, orig_wg = T.bn_relu_bwd, T.linear_wgrad
def bwd(dz, argmax, ns_, y, Cc, *a):
    dy, s = orig_bwd(dz, argmax, ns_, y, Cc, *a)
    rec.setdefault("bwd", []).append((dz.clone(), None if argmax is None else argmax.clone(), y.clone(), dy.clone(), s.clone()))
    rec.setdefault("vecs", []).append([t.clone() for t in a])
    dy2, s2 = orig_bwd(dz.contiguous(), argmax, 1 if argmax is None else ns_, y, Cc, *a)
    print("   [bwd call] C", Cc, "ns", ns_, "dz", tuple(dz.shape), dz.stride(), "y", tuple(y.shape), y.stride(), "repeat diff %.2e" % float((dy2-dy).abs().max()))
    return dy, s
T.bn_relu_bwd = bwd
mod = modules.PointnetSAModuleVotes(mlp=list(mlp), radius=radius, nsample=ns, normalize_xyz=True, sample_method="fps")
mod.load_state_dict(sd); mod = mod.to(DEV).train()
f = feats.clone().requires_grad_(True)
new_xyz, new_feats, inds = mod(xyz, f, M)
w = torch.from_numpy(synth.features(tuple(new_feats.shape), seed=7)).to(DEV)
(new_feats * w).sum().backward()
# manual torch chain on the same grouped rows
idx = ops.ball_query(new_xyz.detach(), xyz, radius, ns)
x0 = T.sa_group_rows(xyz, ops.cm_to_pm(feats), new_xyz.detach(), idx, C, radius, True)
R = x0.shape[0]
W0 = sd["mlp_module.layer0.conv.weight"].reshape(128,131).to(DEV); W1 = sd["mlp_module.layer1.conv.weight"].reshape(128,128).to(DEV)
g0,b0 = sd["mlp_module.layer0.normlayer.bn.weight"].to(DEV), sd["mlp_module.layer0.normlayer.bn.bias"].to(DEV)
g1,b1 = sd["mlp_module.layer1.normlayer.bn.weight"].to(DEV), sd["mlp_module.layer1.normlayer.bn.bias"].to(DEV)
xt = x0[:, :131].clone().requires_grad_(True); W0t = W0.clone().requires_grad_(True)
y0 = xt @ W0t.t(); y0.retain_grad()
z0 = torch.relu(F.batch_norm(y0.t().reshape(1,128,R), None, None, g0, b0, training=True)[0].t()); z0.retain_grad()
y1 = z0 @ W1.t(); y1.retain_grad()
z1 = torch.relu(F.batch_norm(y1.t().reshape(1,128,R), None, None, g1, b1, training=True)[0].t())
out = z1.reshape(B*M, ns, 128).max(1)[0]
rel(new_feats.permute(0,2,1).reshape(B*M,128), out, "forward vs manual")
dout = w.permute(0,2,1).reshape(B*M,128).contiguous()
out.backward(dout)
(dzL, argL, yL, dyL, sL), (dz0r, arg0, y0r, dy0r, s0r) = rec["bwd"]
rel(dzL, dout, "dz given to last layer"); rel(yL, y1, "y1 saved"); rel(dyL, y1.grad, "dy1")
rel(dz0r, z0.grad, "dz0"); rel(y0r, y0, "y0 saved"); rel(dy0r, y0.grad, "dy0")
rel(mod.mlp_module.layer0.conv.weight.grad.reshape(128,131), W0t.grad, "dW0 (module)")

ka0,kb0,m0,r0 = T.bn_train_finalize(T.col_stats(y0r, 128), R, g0, b0, 1e-5, 0.1, None, None)
for name, a, b in zip(("ka","kb","mean","rstd","gamma"), rec["vecs"][1], (ka0,kb0,m0,r0,g0)):
    rel(a, b, "layer0 " + name)
dyf, sf = orig_bwd(dz0r, None, 1, y0r, 128, ka0, kb0, m0, r0, g0)
rel(dyf, y0.grad, "dy0 with fresh vectors")
