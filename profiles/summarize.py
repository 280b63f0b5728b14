#!/usr/bin/env python
"""Summaries of ncu outputs for profiles/ (run here, no GPU needed).

    python profiles/summarize.py launches gpurun_out/<launches>.csv          # per-kernel time shares of a launch list
    python profiles/summarize.py raw gpurun_out/<capture>.ncu-rep            # key metrics per captured launch
    python profiles/summarize.py hot gpurun_out/<capture>.ncu-rep [n]        # hottest SASS lines of the first kernel
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path, per_launch=False):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    rows = []
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")[:48]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        rows.append((name, row["Grid Size"], v))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("%-48s %5s %11s %7s" % ("kernel", "n", "total us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-48s %5d %11.1f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
    print("%-48s %5d %11.1f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))
    if per_launch:
        for r in rows:
            print("%-48s %-16s %9.1f us" % r)


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__inst_executed.sum"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== %s  grid %s block %s" % (d.get("Kernel Name", "?")[:70], d.get("Grid Size"), d.get("Block Size")))
        for k in KEYS:
            if k in d:
                print("   %-72s %s %s" % (k, d[k], units[hdr.index(k)]))


def hot(path, n=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, block = None, []
    name = None
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            if block:
                break
            name = r[1]
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            block.append(dict(zip(hdr, r)))
    tot = sum(int(x["# Samples"] or 0) for x in block) or 1
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print(name, "| samples", tot)
    agg = collections.Counter()
    for x in block:
        for k in stalls:
            agg[k] += int(x[k] or 0)
    print("stall totals:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in agg.most_common(8)))
    for x in sorted(block, key=lambda x: -int(x["# Samples"] or 0))[:n]:
        s = {k: int(x[k] or 0) for k in stalls}
        big = max(s.items(), key=lambda kv: kv[1])
        print("%6.2f%%  %-72s %s" % (100 * int(x["# Samples"]) / tot, x["Source"][:72], big[0]))


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2], per_launch=len(sys.argv) > 3)
    elif cmd == "raw":
        raw(sys.argv[2])
    else:
        hot(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
