#!/usr/bin/env python
"""SASS of one kernel of a built library, one instruction per line with the source line it belongs to, plus counts of
control instructions -- the offline check used while flattening the decoder warp's loop (no GPU needed).

  python tools/sass_of.py <lib.so> <kernel-substring> [file-filter] > listing.txt
"""
import os, re, subprocess, sys, tempfile

lib, kname = sys.argv[1], sys.argv[2]
ffilter = sys.argv[3] if len(sys.argv) > 3 else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
func, cur, inl = None, None, ""
ctrl = ("BRA", "BSSY", "BSYNC", "CALL", "RET", "WARPSYNC", "BRX", "JMP", "EXIT", "BREAK")
n_ins = n_ctrl = 0
for ln in dis:
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
    if m:
        func, cur = m.group(1), None
        continue
    if func is None or kname not in func:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = "%s:%s" % (m.group(1).split('/')[-1], m.group(2))
        continue
    m = re.match(r'\s*(\.L_x_\d+):', ln)
    if m:
        print(m.group(1) + ":")
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        txt = m.group(2).strip()
        op = txt.split()
        name = op[1] if op[0].startswith('@') and len(op) > 1 else op[0]
        isc = name.split('.')[0] in ctrl
        n_ins += 1
        n_ctrl += isc
        if ffilter is None or (cur and ffilter in cur):
            print("  %05x %s %-70s ; %s" % (int(m.group(1), 16), "*" if isc else " ", txt[:70], cur or ""))
print("# %d instructions, %d control" % (n_ins, n_ctrl))
