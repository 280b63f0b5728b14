#!/usr/bin/env python
"""BASELINE.json configs[2] and configs[4] bench lines: the Pedestrian / sparse regime (N = 512 / 512,
batch 128) and the N x batch sweep of the synthetic crops, one `bench.py` line each, appended to a .jsonl file.

    python tools/sweep.py out.jsonl [--quick]
"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = sys.argv[1]
quick = "--quick" in sys.argv
runs = [dict(name="config3 pedestrian sparse", args=["--batch", "128", "--nsearch", "512", "--ntemplate", "512", "--kind", "sparse"])]
for n in (256, 512, 1024, 2048):
    for b in ((16, 128) if quick else (16, 64, 256)):
        runs.append(dict(name="config5 N=%d B=%d" % (n, b), args=["--batch", str(b), "--nsearch", str(n), "--ntemplate", str(n // 2)]))
common = ["--no-cpu-baseline", "--no-reference-gpu", "--no-whole-model", "--no-tracking", "--no-train-step", "--sustain-s", "1.0"]
with open(out, "w") as f:
    for r in runs:
        p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + r["args"] + common, capture_output=True, text=True)
        line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""
        try:
            d = json.loads(line)
            d["sweep_case"] = r["name"]
            keep = {k: d[k] for k in ("sweep_case", "metric", "value", "unit", "ms_per_step", "n_gpus", "config", "e2e", "sequential_l2_flushed",
                                      "sustained", "roofline", "stage_ms", "stage_roofline", "clocks", "gpu_launches") if k in d}
            f.write(json.dumps(keep) + "\n")
            print(r["name"], round(d["value"]), "frames/s", round(d["ms_per_step"], 3), "ms; top", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
        except Exception as e:
            f.write(json.dumps({"sweep_case": r["name"], "error": p.stderr[-500:]}) + "\n")
            print(r["name"], "FAILED", p.stderr[-300:])
        f.flush()
