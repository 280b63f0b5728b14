#!/usr/bin/env python
"""Latency of the whole tracker forward (HotPath.forward_full, CUDA-graph replay) at small batches -- the regime of the
reference's evaluation loop, which tracks ONE tracklet frame by frame at batch 1 (tools/test_tracking.py,
eval_tracking_utils.py:140-274).  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ptt_b200 import hotpath, synth

hp = hotpath.HotPath(synth.full_model_state_dict(0), device="cuda:0")
out = {}
for B in (1, 2, 4, 8, 16, 48):
    s = torch.from_numpy(synth.make_clouds(B, 1024, 10 + B, "dense")).cuda()
    t = torch.from_numpy(synth.make_clouds(B, 512, 20 + B, "dense", role="template")).cuda()
    for _ in range(5):
        hp.forward_graph(s, t, full=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 100
    e0.record()
    for _ in range(n):
        hp.forward_graph(s, t, full=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out["B=%d" % B] = {"ms_per_forward": round(ms, 4), "frames_per_s": round(B / ms * 1e3, 1)}
print(json.dumps({"what": "HotPath.forward_full, back-to-back CUDA-graph replays, inputs resident (N = 1024 / 512)", **out}))
