#!/usr/bin/env python
"""Executed warp instructions per symbol by source line for one kernel of an ncu report (--set full --import-source on),
split into the decoder warp's loop (lcv_decode_stream) and the rest (updater warp, setup).

  python tools/dec_lines.py <report.ncu-rep> [kernel-substring] [symbols] [top]
Needs the library the report was captured with (image_compression_2_b200/liblatentcodec.so, built with -lineinfo)."""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
kname = sys.argv[2] if len(sys.argv) > 2 else "lc_decode_v2_w8_kernel"
nsym = float(sys.argv[3]) if len(sys.argv) > 3 else 1024 * 8192.0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
opfilter = sys.argv[5] if len(sys.argv) > 5 else None  # e.g. "MOV|BSSY|BSYNC": count only these opcodes
lib = os.path.join(ROOT, "image_compression_2_b200", "liblatentcodec.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
func, cur, table = None, None, collections.defaultdict(list)
for ln in dis:
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
    if m:
        func, cur = m.group(1), None
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m and func:
        table[func].append((int(m.group(1), 16), m.group(2).strip(), cur))
ins = table[[f for f in table if kname in f][0]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kname], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
data = [r for r in rows[2:] if len(r) == len(h) and r[0] != "Address"][:len(ins)]
ii, si = h.index("Instructions Executed"), h.index("# Samples")
assert len(data) == len(ins), (len(data), len(ins))
src = {}


def text(f, ln):
    p = os.path.join(ROOT, "image_compression_2_b200", "csrc", f)
    if os.path.exists(p):
        if p not in src:
            src[p] = open(p).read().split("\n")
        return src[p][ln - 1].strip()[:100]
    return ""


# the decoder warp's loop: the contiguous address range of lines inside lcv_decode_stream
v2 = open(os.path.join(ROOT, "image_compression_2_b200", "csrc", "lc_decoder_v2.cuh")).read().split("\n")
l0 = next(i for i, t in enumerate(v2) if "void lcv_decode_stream(" in t) + 1
l1 = next(i for i, t in enumerate(v2) if "// Block entry:" in t) + 1
addrs = [a for a, t, c in ins if c and c[0] == "lc_decoder_v2.cuh" and l0 <= c[1] < l1]
if os.environ.get("DEC_RANGE") == "all" or not addrs:  # any other kernel: one range
    addrs = [a for a, t, c in ins]
lo, hi = min(addrs), max(addrs)
tot_s = sum(float(r[si] or 0) for r in data)
agg = {True: collections.defaultdict(lambda: [0.0, 0.0]), False: collections.defaultdict(lambda: [0.0, 0.0])}
ops = collections.Counter()
for k, (a, t, c) in enumerate(ins):
    d = lo <= a <= hi
    if opfilter and not re.search(opfilter, t):
        continue
    x, s = float(data[k][ii] or 0) / nsym, float(data[k][si] or 0)
    agg[d][c or ("?", 0)][0] += x
    agg[d][c or ("?", 0)][1] += s
    if d:
        ops[t.split()[1] if t.startswith("@") else t.split()[0]] += x
for d, name in ((True, "decoder warp loop"), (False, "updater warp and setup")):
    ti = sum(v[0] for v in agg[d].values())
    ts = sum(v[1] for v in agg[d].values())
    print("== %s: %.1f warp instructions/symbol, %.1f%% of the stall samples" % (name, ti, 100 * ts / tot_s))
    for key, v in sorted(agg[d].items(), key=lambda x: -x[1][0])[:top if d else top // 3]:
        print("%6.2f inst/sym %5.1f%% samples  %s:%d  %s" % (v[0], 100 * v[1] / tot_s, key[0], key[1], text(*key)))
print("== decoder warp loop by opcode")
print(", ".join("%s %.1f" % (o, x) for o, x in ops.most_common(30)))
if os.environ.get("DEC_LINES_SASS"):
    with open(os.environ["DEC_LINES_SASS"], "w") as f:
        for k, (a, t, c) in enumerate(ins):
            if lo <= a <= hi:
                f.write("%04x %5.2f %6s  %-22s %s\n" % (a, float(data[k][ii] or 0) / nsym, data[k][si], "%s:%d" % (c[0][-16:], c[1]) if c else "?", t))
