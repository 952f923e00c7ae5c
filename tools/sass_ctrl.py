#!/usr/bin/env python
"""SASS of one kernel with the scheduling control fields decoded from the instruction encoding (Volta+ layout of the
upper 64-bit word: stall count bits 41-44, yield 45, write barrier 46-48, read barrier 49-51, wait mask 52-57):
which scoreboard slot every variable-latency instruction sets and which slots every instruction waits for.

  python tools/sass_ctrl.py <lib.so> <kernel-substring> [lo_hex hi_hex]
"""
import os, re, subprocess, sys, tempfile

lib, kname = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
names = subprocess.run(["cuobjdump", "-elf", cubin], capture_output=True, text=True).stdout
fun = sorted(set(re.findall(r'\.text\.(\S*%s\S*)' % re.escape(kname), names)), key=len)[0]
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, cubin], capture_output=True, text=True).stdout.split("\n")
i = 0
while i < len(out):
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/', out[i])
    if m and i + 1 < len(out):
        m2 = re.match(r'\s*/\* 0x([0-9a-f]{16}) \*/', out[i + 1])
        a = int(m.group(1), 16)
        if m2 and lo <= a <= hi:
            w = int(m2.group(1), 16)
            stall, yld = (w >> 41) & 15, (w >> 45) & 1
            wb, rb, wm = (w >> 46) & 7, (w >> 49) & 7, (w >> 52) & 63
            print("%05x st=%2d %s w%s r%s wait=%-6s %s" % (a, stall, "Y" if yld else "-", wb if wb != 7 else "-",
                                                        rb if rb != 7 else "-", "".join(str(b) for b in range(6) if wm >> b & 1) or "-",
                                                        m.group(2).strip()))
        i += 2
        continue
    i += 1
