"""Tuning aid: clock64 timeline of CTA 0 of the fused SA kernel (SA2-like or SA1-like layer at B = 48)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptt_b200 import _lib, hotpath, ops, synth
which = sys.argv[1] if len(sys.argv) > 1 else "sa2"
sd = synth.hot_path_state_dict(0)
hp = hotpath.HotPath(sd)
L = _lib.lib()
L.ptt_debug_sa_timeline.argtypes = [ctypes.c_void_p]
B = 48
if which == "sa1":
    packed, N, M, C, r = hp.sa[0], 1024, 512, 0, 0.3
else:
    packed, N, M, C, r = hp.sa[1], 512, 256, 128, 0.5
xyz = torch.from_numpy(synth.make_clouds(B, N, 5, "dense")).cuda()
feats = torch.randn(B, N, C, device="cuda") if C else None
inds, new_xyz = ops.furthest_point_sampling(xyz, M, return_new_xyz=True)
idx = ops.ball_query(new_xyz, xyz, r, 32)
dbg = torch.zeros(5000, dtype=torch.int64, device="cuda")
for rep in range(3):
    dbg.zero_()
    L.ptt_debug_sa_timeline(dbg.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.sa_mlp_fwd(packed, xyz, feats, new_xyz, idx, r, True, want_pm=True, want_cm=False)
    e1.record()
    torch.cuda.synchronize()
L.ptt_debug_sa_timeline(None)
print(which, "stage ms (incl. G' contraction):", e0.elapsed_time(e1))
d = dbg.cpu().numpy()
t0 = int(d[1])
names = {1: "wait ha_full", 2: "G2 issue", 3: "wait hb_full", 4: "G3 issue", 5: "tile end"}
prev = t0
for i in range(60):
    tag, t = int(d[2 * i]), int(d[2 * i + 1])
    if tag <= 0:
        break
    print("%-14s %8d (+%d)" % (names[tag], t - t0, t - prev))
    prev = t
print("epilogue (d2_full, hb_full sent, d3_full, E3 done):")
for it in range(6):
    print("  tile", it, [int(d[1000 + it * 4 + j]) - t0 for j in range(4)])
print("producer (start wait ha_free, got it, H1 written):")
for it in range(6):
    print("  tile", it, [int(d[2000 + it * 3 + j]) - t0 for j in range(3)])
print("producer loop top (enter, after bar 1, after s_info written, after bar 2):")
for it in range(6):
    print("  tile", it, [int(d[3000 + it * 4 + j]) - t0 for j in (0, 1, 3, 2)])
