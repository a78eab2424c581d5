#!/usr/bin/env python
"""Per-CUDA-source-line instruction counts from an ncu report (built with -lineinfo).

usage: python profiles/line_profile.py <report.ncu-rep> <lib.so> <kernel-substring> [top]
Joins `ncu --page source --csv` (SASS rows with 'Instructions Executed' and stall samples) with
`nvdisasm -g` line markers of the same kernel, by instruction order.
"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
lines = []
for f in sorted(os.listdir(td)):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, f)], capture_output=True, text=True).stdout
    # split per function
    cur, active, cur_line = None, False, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            active = kern in m.group(1) and not lines
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur_line)
    if lines:
        break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[hi + 1:] if len(r) > ie]
assert abs(len(body) - len(lines)) < 8, (len(body), len(lines))
agg, smp = defaultdict(int), defaultdict(int)
tot = tots = 0
for r, l in zip(body, lines):
    n, s = int(r[ie]), int(r[ss])
    agg[l] += n; smp[l] += s; tot += n; tots += s
print("total warp-instructions", tot, "samples", tots, "sass", len(body))
srcs = {}
for (f, l), n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    if f not in srcs:
        for root, _, fs in os.walk(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "jsplayer_b200")):
            if f in fs:
                srcs[f] = open(os.path.join(root, f)).read().splitlines()
    text = srcs.get(f, [""] * (l + 1))[l - 1].strip()[:100] if f in srcs else ""
    print("%5.1f%% inst %5.1f%% stall  %s:%d  %s" % (100.0 * n / tot, 100.0 * smp[(f, l)] / max(1, tots), f, l, text))
