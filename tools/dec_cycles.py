#!/usr/bin/env python
"""Where the decoder warp's cycles go, instruction by instruction, from one ncu report (--set full --import-source on).

The decoder warp of a block is resident (and on its serial chain) for the whole kernel, so its stall samples are a
time profile of that chain: samples(instruction) / samples(decoder loop) x cycles-per-symbol = cycles that instruction
costs per symbol; divided by how often it executes per symbol = cycles per execution.

  python tools/dec_cycles.py <report.ncu-rep> <lib.so> [kernel] [streams] [ms]  > listing
Prints the loop's instructions in address order:  address, executions/symbol, cycles/execution, cycles/symbol, SASS, line.
"""
import csv, io, os, re, subprocess, sys, tempfile

rep, lib = sys.argv[1], sys.argv[2]
kname = sys.argv[3] if len(sys.argv) > 3 else "lc_decode_v2_w8_kernel"
streams = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
nsym = streams * 8192.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kname], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
data = [r for r in rows[2:] if len(r) == len(h) and r[0] != "Address"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", kname], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
dur_ns = float(rr[2][rr[0].index("gpu__time_duration.sum")].replace(",", "")) if "gpu__time_duration.sum" in rr[0] else 0.0
unit = rr[1][rr[0].index("gpu__time_duration.sum")] if dur_ns else ""
ms = float(sys.argv[5]) if len(sys.argv) > 5 else (dur_ns / 1e6 if unit in ("ns", "nsecond") else dur_ns if unit in ("ms", "msecond") else dur_ns / 1e3)
cyc_sym = ms * 1e-3 * 1.965e9 / 8192.0  # cycles per symbol of one stream (one wave: every stream runs the whole kernel)
# line info from the library
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
func, cur, lines = None, None, {}
for ln in dis:
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
    if m:
        func, cur = m.group(1), None
        continue
    if func is None or kname not in func:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = "%s:%s" % (m.group(1).split('/')[-1], m.group(2))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        lines[int(m.group(1), 16)] = cur
base = int(data[0][0], 16)
ie, si = h.index("Instructions Executed"), h.index("# Samples")
reasons = [c for c in h if c.startswith("stall_") and "(Not Issued)" not in c]
# the decoder warp's loop: instructions whose source line lies inside the per-symbol lambda of the decoder-stream function
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spans = {}
for fn in ("lc_decoder_v2_flat.cuh", "lc_decoder_v2.cuh"):
    src = open(os.path.join(ROOT, "image_compression_2_b200", "csrc", fn)).read().split("\n")
    a = [i + 1 for i, s in enumerate(src) if "auto one_symbol = " in s]
    b = [i + 1 for i, s in enumerate(src) if "auto row_done = " in s]
    if a and b:
        spans[fn] = (a[0], b[0])
def in_loop(a):
    c = lines.get(a) or ""
    f, _, l = c.partition(":")
    return f in spans and l.isdigit() and spans[f][0] <= int(l) < spans[f][1]
addrs = [int(r[0], 16) - base for r in data]
loop_a = [a for a in addrs if in_loop(a)]
lo, hi = (min(loop_a), max(loop_a)) if loop_a else (0, 1 << 30)
loop = [r for r in data if lo <= int(r[0], 16) - base <= hi]
tot = sum(float(r[si]) for r in loop)
print("# kernel %.3f ms, %.0f cycles/symbol/stream; decoder loop %05x..%05x: %.1f warp instructions/symbol, %.1f%% of the kernel's samples"
      % (ms, cyc_sym, lo, hi, sum(float(r[ie]) for r in loop) / nsym, 100 * tot / sum(float(r[si]) for r in data)))
for name in sorted(reasons, key=lambda n: -sum(float(r[h.index(n)] or 0) for r in loop))[:9]:
    t = sum(float(r[h.index(name)] or 0) for r in loop)
    print("#   %-26s %5.1f%%  %5.0f cycles/symbol" % (name, 100 * t / tot, cyc_sym * t / tot))
for r in loop:
    a = int(r[0], 16) - base
    ex = float(r[ie]) / nsym
    if ex < 0.004:
        continue
    cs = cyc_sym * float(r[si]) / tot
    top = sorted(((float(r[h.index(n)] or 0), n[6:]) for n in reasons), reverse=True)[0]
    print("%05x %5.2f x %6.1f = %5.1f  %-60s ; %s %s" % (a, ex, cs / ex, cs, r[1].strip()[:60], lines.get(a) or "",
                                                        top[1] if cs >= 4 else ""))
