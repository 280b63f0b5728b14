"""Tuning aid: time every (threads, points-per-thread) variant of the FPS kernel for the live shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptt_b200 import _lib, synth

L = _lib.lib()
fn = L.ptt_fps_variant
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
               ctypes.c_void_p]
B = 48
for N, M in ((1024, 512), (512, 256), (128, 64), (2048, 1024), (256, 128)):
    xyz = torch.from_numpy(synth.make_clouds(B, N, 3, "dense")).cuda()
    idx = torch.empty(B, M, dtype=torch.int32, device="cuda")
    new_xyz = torch.empty(B, M, 3, device="cuda")
    res = []
    for threads in (32, 64, 128, 256, 512):
        for ppt in (1, 2, 4, 8, 16):
            if threads * ppt < N:
                continue
            st = torch.cuda.current_stream().cuda_stream
            if fn(xyz.data_ptr(), B, N, M, idx.data_ptr(), new_xyz.data_ptr(), threads, ppt, st) != 0:
                continue
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn(xyz.data_ptr(), B, N, M, idx.data_ptr(), new_xyz.data_ptr(), threads, ppt, st)
            e1.record()
            torch.cuda.synchronize()
            res.append((e0.elapsed_time(e1) / 10 * 1e3, threads, ppt))
    res.sort()
    print("N=%d M=%d:" % (N, M), ", ".join("%dx%d %.1fus" % (t, p, us) for us, t, p in res[:6]))
