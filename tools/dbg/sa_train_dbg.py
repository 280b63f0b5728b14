import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from ptt_b200 import modules, synth
from test_oracle_golden import sa_state_dict
DEV="cuda:0"
def run(B,N,cin,mlp,npoint,radius,ns):
    sd = {k: torch.from_numpy(v) for k, v in synth.fill_state_dict(sa_state_dict(mlp), seed=500 + N).items()}
    xyz = torch.from_numpy(synth.make_clouds(B, N, 600 + N, "dense", role="template")).to(DEV)
    feats = torch.from_numpy(synth.features((B, cin, N), seed=601 + N)).to(DEV)
    outs=[]
    for native in (True, False):
        mod = modules.PointnetSAModuleVotes(mlp=list(mlp), radius=radius, nsample=ns, normalize_xyz=True, sample_method="fps")
        mod.load_state_dict(sd); mod = mod.to(DEV).train(); mod.native_train = native
        f = feats.clone().requires_grad_(True)
        new_xyz, new_feats, inds = mod(xyz, f, npoint)
        w = torch.from_numpy(synth.features(tuple(new_feats.shape), seed=7)).to(DEV)
        (new_feats * w).sum().backward()
        outs.append(dict(fg=f.grad, **{k: p.grad for k, p in mod.named_parameters()}))
    a,b = outs
    for k in b:
        sc = float(b[k].abs().max()); err = float((a[k]-b[k]).abs().max())
        print("  %-45s err %.3e scale %.3e rel %.2e" % (k, err, sc, err/max(sc,1e-9)))
for mlp in ([128,128],[128,128,128],[128,128,128,256]):
    print("mlp", mlp); run(4,512,128,mlp,256,0.5,32)
print("ns=1"); run(4,512,128,[128,128,128],256,0.5,1)
