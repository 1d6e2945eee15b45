#!/usr/bin/env python
"""Per-source-line instruction counts of a kernel from an ncu report (needs -lineinfo).
usage: ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie = [i for i, h in enumerate(hdr) if h == "Instructions Executed"][0]
        ss = [i for i, h in enumerate(hdr) if h == "# Samples"][0]
        continue
    if hdr is None or r[0] in ("Function Name", "Kernel Name"):
        continue
    if r[0] != "":  # a source line aggregate
        try:
            agg[(cur_file, int(r[0]), r[1].strip()[:90])] = (int(r[ie]), int(r[ss]))
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print(f"total warp instructions {tot}, samples {tots}")
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/tot:5.1f}% inst {100*s/tots:5.1f}% stall  {f}:{ln}: {src}")
