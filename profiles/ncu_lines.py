#!/usr/bin/env python
"""Per-source-line instruction / stall summary of one .ncu-rep (needs -lineinfo + --import-source on).

    python profiles/ncu_lines.py gpurun_out/x.ncu-rep [min_percent] [kernel-id]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, hdr, kernel = None, None, None
items = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        kernel = r[1][:60]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or r[0] in ("", "Kernel Name"):
        continue
    try:
        d = dict(zip(hdr[:2] + hdr[4:], r[:2] + r[4:]))      # the aggregated line row: Line No, Source, then metrics
        n = int(d["Instructions Executed"]); smp = int(d["Warp Stall Sampling (All Samples)"])
    except Exception:
        continue
    items.append((kernel, cur_file, r[0], n, smp, r[1]))
kern = sorted(set(i[0] for i in items))
for k in kern:
    it = [i for i in items if i[0] == k]
    tot = sum(i[3] for i in it) or 1; tots = sum(i[4] for i in it) or 1
    print(f"== {k}: {tot} warp instructions, {tots} samples")
    for (_, f, ln, n, smp, src) in it:
        if 100.0 * n / tot >= minp or 100.0 * smp / tots >= minp:
            print(f"{f:>16}:{ln:>4} inst {100.0 * n / tot:5.1f}%  samples {100.0 * smp / tots:5.1f}% | {src.strip()[:110]}")
