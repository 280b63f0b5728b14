"""Accuracy report: max |error| of every hot-path output against the CPU oracle, per code path."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import torch_port
from ptt_b200 import _lib, hotpath, synth

L = _lib.lib()
L.ptt_debug_set_cluster.argtypes = [ctypes.c_int]
L.ptt_debug_force_ffma.argtypes = [ctypes.c_int]
sd = synth.hot_path_state_dict(0)
cases = {"smoke": (dict(npoints_search=(128, 64, 32), npoints_template=(64, 32, 16), box_npoint=16), 256, 128),
         "yaml": (None, 1024, 512)}
for cname, (cfg, ns, nt) in cases.items():
    search = synth.make_clouds(2, ns, 11, "dense")
    template = synth.make_clouds(2, nt, 12, "dense", role="template")
    want = torch_port.hot_path_frame(sd, torch.from_numpy(search), torch.from_numpy(template), cfg)
    for mode, (cl, ff) in {"default (tcgen05 fused)": (0, 0), "generic transformer path, tcgen05 GEMMs": (-1, 0),
                           "generic path, CUDA-core fp32 GEMMs": (-1, 1)}.items():
        L.ptt_debug_set_cluster(cl); L.ptt_debug_force_ffma(ff)
        hp = hotpath.HotPath(sd, cfg=cfg)
        out = hp(torch.from_numpy(search).cuda(), torch.from_numpy(template).cuda())
        torch.cuda.synchronize()
        L.ptt_debug_set_cluster(0); L.ptt_debug_force_ffma(0)
        row = []
        for k in ("search_feats", "template_feats", "centroid_feats", "box_sa_feats", "box_feats"):
            a, b = out[k].cpu().numpy().astype(np.float64), want[k].numpy().astype(np.float64)
            err = np.abs(a - b)
            viol = (err / (1e-4 + 1e-4 * np.abs(b))).max()
            row.append("%s %.1e (%.2f of tol, |x|max %.1f)" % (k.split("_")[0], err.max(), viol, np.abs(b).max()))
        print("%-6s %-42s %s" % (cname, mode, "; ".join(row)))
