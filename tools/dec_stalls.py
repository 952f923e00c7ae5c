#!/usr/bin/env python
"""Stall samples of the decoder warp's loop by reason, with the instructions that collect them.
  python tools/dec_stalls.py <report.ncu-rep> [kernel-substring] [top]   (see tools/dec_lines.py)"""
import csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
kname = sys.argv[2] if len(sys.argv) > 2 else "lc_decode_v2_w8_kernel"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kname], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
data = [r for r in rows[2:] if len(r) == len(h) and r[0] != "Address"]
base = int(data[0][0], 16)
s = h.index("# Samples")
tot = sum(float(r[s]) for r in data)
# the loop = from the first to the last instruction executed about once per symbol (8192 symbols x streams)
ie = h.index("Instructions Executed")
mx = max(float(r[ie]) for r in data)
per_sym = [i for i, r in enumerate(data) if 0.2 * mx / 20 < float(r[ie])]  # coarse; refine below
reasons = [c for c in h if c.startswith("stall_") and "(Not Issued)" not in c]
lo = int(os.environ.get("DEC_LO", "0"), 16) if os.environ.get("DEC_LO") else None
hi = int(os.environ.get("DEC_HI", "0"), 16) if os.environ.get("DEC_HI") else None
D = [r for r in data if lo is None or lo <= int(r[0], 16) - base <= hi]
print("samples in range: %.1f%% of the kernel's" % (100 * sum(float(r[s]) for r in D) / tot))
for name in sorted(reasons, key=lambda n: -sum(float(r[h.index(n)] or 0) for r in D)):
    i = h.index(name)
    t = sum(float(r[i] or 0) for r in D)
    if t < 0.003 * tot:
        continue
    print("== %s: %.1f%% of all samples" % (name, 100 * t / tot))
    for r in sorted(D, key=lambda r: -float(r[i] or 0))[:top]:
        print("   %04x %6s  %s" % (int(r[0], 16) - base, r[i], r[1].strip()[:90]))
