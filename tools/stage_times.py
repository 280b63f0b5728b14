"""Tuning aid: per-step, per-stage CUDA-event times of the eager single-stream hot path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptt_b200 import hotpath, synth
hp = hotpath.HotPath(synth.hot_path_state_dict(0))
s = torch.from_numpy(synth.make_clouds(48, 1024, 1, "dense")).cuda()
t = torch.from_numpy(synth.make_clouds(48, 512, 2, "dense", role="template")).cuda()
for _ in range(3):
    hp(s, t)
torch.cuda.synchronize()
hp.overlap = False
hp.profile(True)
for _ in range(6):
    hp(s, t)
torch.cuda.synchronize()
for k, v in sorted(hp.stage_events.items()):
    print("%-28s" % k, " ".join("%7.1f" % (a.elapsed_time(b) * 1e3) for a, b in v))
