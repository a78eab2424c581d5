#!/usr/bin/env python
"""Aggregates profiles/line_profile.py output (instruction and stall-sample shares per source line) by code region.
usage: python profiles/region_profile.py <report.ncu-rep> <lib.so> <kernel-substring>"""
import re, subprocess, sys, os
here = os.path.dirname(os.path.abspath(__file__))
out = subprocess.run([sys.executable, os.path.join(here, "line_profile.py"), sys.argv[1], sys.argv[2], sys.argv[3], "100000"],
                     capture_output=True, text=True).stdout
src = {}
def region(f, l):
    """names the enclosing function of line l in file f (nearest preceding '__device__' / '__global__' line)"""
    if f not in src:
        for root, _, fs in os.walk(os.path.join(here, "..", "jsplayer_b200")):
            if f in fs:
                src[f] = open(os.path.join(root, f)).read().splitlines()
        src.setdefault(f, [])
    L = src[f]
    for i in range(min(l, len(L)) - 1, -1, -1):
        m = re.search(r"__(?:device|global)__.*?(\w+)\s*\(", L[i])
        if m and not L[i].lstrip().startswith("//"):
            return "%s:%s" % (f, m.group(1))
    return f
agg = {}
print(out.splitlines()[0])
for ln in out.splitlines():
    m = re.match(r"\s*([\d.]+)% inst\s+([\d.]+)% stall\s+(\S+):(\d+)", ln)
    if not m:
        continue
    a = agg.setdefault(region(m.group(3), int(m.group(4))), [0.0, 0.0])
    a[0] += float(m.group(1)); a[1] += float(m.group(2))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[0] >= 0.3 or v[1] >= 0.3:
        print("%-48s inst %5.1f%%  stall %5.1f%%" % (k, v[0], v[1]))
