#!/usr/bin/env python
"""Stall-reason totals and a per-region (N SASS instructions) sample histogram of the first kernel in an ncu report.
usage: ncu_regions.py report.ncu-rep [chunk] ; also writes the annotated SASS to /tmp/ncu_sass.txt"""
import csv, subprocess, collections, sys
rep = sys.argv[1]
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 200
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; stall = collections.Counter(); lines = []; nk = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        nk += 1
        continue
    if nk > 1:
        break
    if r[0] == "Address":
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None:
        continue
    st = {h[6:]: int(r[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i] not in ("", "0")}
    for k, v in st.items():
        stall[k] += v
    lines.append((int(r[ix['# Samples']]), int(r[ix['Instructions Executed']]), r[1].strip(), sorted(st.items(), key=lambda kv: -kv[1])[:2]))
tot = sum(stall.values())
print("total samples", tot, "static instructions", len(lines), "dynamic warp instructions", sum(l[1] for l in lines))
for h, v in stall.most_common(10):
    print(f"  {h:22s} {v:7d} {100*v/tot:5.1f}%")
for i in range(0, len(lines), chunk):
    s = sum(l[0] for l in lines[i:i + chunk]); n = sum(l[1] for l in lines[i:i + chunk])
    ops = collections.Counter((l[2].split()[1] if l[2].startswith('@') else l[2].split()[0]).split('.')[0] for l in lines[i:i + chunk])
    print(f"  [{i:5d}] samples {s:6d} ({100*s/tot:4.1f}%) dyn-inst {n:10d}  {ops.most_common(4)}")
open('/tmp/ncu_sass.txt', 'w').write("\n".join(f"{a:6d} {b:9d} {c:70s} {d}" for a, b, c, d in lines))
