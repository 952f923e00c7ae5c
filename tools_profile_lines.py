#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS listing with nvdisasm line info: per-source-line samples/instructions.
usage: tools_profile_lines.py <src.csv> <dis.txt> <kernel-substr> [top]"""
import csv, re, sys, collections
src_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
# nvdisasm: collect per-function instruction list with current (file,line), honoring inlined-at
lines = open(dis).read().splitlines()
func = None; cur = None; table = collections.defaultdict(list)
for ln in lines:
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
    if m: func = m.group(1); cur = None; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), m.group(3)); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m and func: table[func].append((int(m.group(1), 16), m.group(2).strip(), cur))
fn = [f for f in table if kname in f][0]
ins = table[fn]
rows = list(csv.reader(open(src_csv)))
h = rows[1]
si, ii = h.index('# Samples'), h.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) == len(h)]
print('kernel', fn, 'sass', len(ins), 'csv rows', len(data))
agg = collections.defaultdict(lambda: [0.0, 0.0])
tots = toti = 0.0
n = min(len(ins), len(data))
for k in range(n):
    r = data[k]; s = float(r[si] or 0); i = float(r[ii] or 0)
    key = ins[k][2][:2] if ins[k][2] else ('?', 0)
    agg[key][0] += s; agg[key][1] += i; tots += s; toti += i
print('total samples %.0f, warp instructions %.0f' % (tots, toti))
srccache = {}
def srcline(f, l):
    import glob
    if f not in srccache:
        c = glob.glob('/root/repo/image_compression_2_b200/csrc/' + f)
        srccache[f] = open(c[0]).read().splitlines() if c else []
    L = srccache[f]
    return L[l - 1].strip()[:100] if 0 < l <= len(L) else ''
for (f, l), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100 * s / tots, 100 * i / toti, f, l, srcline(f, l)))
