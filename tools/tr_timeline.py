"""Tuning aid: clock64 timeline of CTA 0 of one transformer pair-row pass (pass 1: generated A, store epilogue)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ptt_b200 import _lib, ops, synth

cs = int(sys.argv[1]) if len(sys.argv) > 1 else 0
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
B, n, k, dm = 48, 128, 16, 512
L = _lib.lib()
L.ptt_debug_set_cluster.argtypes = [ctypes.c_int]
L.ptt_debug_set_cluster(cs)
fn = L.ptt_debug_tr_pass_timeline
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
xyz = torch.from_numpy(synth.make_clouds(B, n, 1, "dense", role="template")).cuda()
knn = ops.knn(xyz, k)
w = torch.randn(dm, dm, device="cuda") / dm ** 0.5
b = torch.randn(dm, device="cuda")
lin = ops.PackedLinear(w, b)
wd0 = torch.randn(4, dm, device="cuda")
out = torch.empty(B * n * k, dm, device="cuda")
aux = torch.randn(B * n, 3 * dm, device="cuda")
big = torch.randn(B * n * k, dm, device="cuda")
dbg = torch.zeros(6000, dtype=torch.int64, device="cuda")
ldw = dm
wimg_ptr = lin.params.data_ptr() + (dm + 1) * ldw * 4
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    dbg.zero_()
    e0.record()
    rc = fn(xyz.data_ptr(), knn.data_ptr(), B, n, k, dm, wd0.data_ptr(), dm, wimg_ptr, b.data_ptr(), out.data_ptr(), dbg.data_ptr(), flags,
            torch.cuda.current_stream().cuda_stream, mode, aux.data_ptr(), big.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0, rc
print("kernel ms:", e0.elapsed_time(e1), "flags", flags, "mode", mode)
d = dbg.cpu().numpy()
t0 = None
ev = []
for i in range(400):
    tag, t = int(d[2 * i]), int(d[2 * i + 1])
    if tag == -1 or (tag == 0 and t == 0):
        break
    if t0 is None:
        t0 = t
    ev.append((tag, t - t0))
print("cluster", cs, "events", len(ev))
if t0 is not None:
    print("producer thread 0: (kb: wait_start, wait_end, produced) rel cycles")
    for i in range(24):
        a0, a1, a2 = (int(d[2000 + i * 3 + j]) for j in range(3))
        if a0:
            print("  unit %d kb %d: %8d %8d %8d   wait %6d produce %6d" % (i // 8, i % 8, a0 - t0, a1 - t0, a2 - t0, a1 - a0, a2 - a1))
prev = 0
for tag, t in ev:
    print("%4d %8d  (+%d)" % (tag, t, t - prev))
    prev = t
if t0 is not None:
    print("epilogue d_full / d_free stamps (rel):", [(int(d[1000 + 2 * i]) - t0, int(d[1001 + 2 * i]) - t0) for i in range(6)])
