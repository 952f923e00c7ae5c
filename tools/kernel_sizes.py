#!/usr/bin/env python
"""SASS size (bytes) of every kernel of a liblatentcodec build: python tools/kernel_sizes.py [lib.so]"""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "image_compression_2_b200", "liblatentcodec.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, last = None, 0
sizes = {}
for ln in out.split("\n"):
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        if name: sizes[name] = last + 16
        name, last = m.group(1), 0
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m: last = int(m.group(1), 16)
if name: sizes[name] = last + 16
for k, v in sorted(sizes.items(), key=lambda kv: -kv[1]):
    d = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
    print("%7d  %s" % (v, d))
