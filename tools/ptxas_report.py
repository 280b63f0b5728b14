"""Register / spill table of every kernel in libptt_b200.so (`nvcc -Xptxas -v` of a forced rebuild).

    python tools/ptxas_report.py [out.txt]

Used before a GPU run to see what a source change did to the hot kernels (spills, register count)."""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def report():
    r = subprocess.run([sys.executable, "-m", "ptt_b200.build", "--force", "-v"], cwd=REPO, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(r.stderr[-4000:])
    rows = []
    name = None
    for line in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name)
            spill = None
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and name:
            spill = (int(m.group(2)), int(m.group(3)))
            continue
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            rows.append((name, int(m.group(1)), spill or (0, 0)))
            name = None
    return sorted(set(rows))


if __name__ == "__main__":
    lines = ["%-110s regs %3d  spill st/ld %4d/%4d" % (n[:110], r, s[0], s[1]) for n, r, s in report()]
    text = "\n".join(lines)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text + "\n")
    print(text)
