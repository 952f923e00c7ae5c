#!/usr/bin/env python
"""Per-state cycle accounting of the fast decoder (debug build with -DLC_DEC_PROFILE).

  python tools/dec_profile.py build     (here: cross-compiles tools/liblatentcodec_prof.so)
  python tools/dec_profile.py [B bits]  (GPU box: runs cfg2-like encode+decode, prints the table)
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIBP = os.path.join(ROOT, "tools", "liblatentcodec_prof.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    from image_compression_2_b200 import build
    print(build.build_library(out=LIBP, defines=("LC_DEC_PROFILE",)))
    sys.exit(0)

os.environ["LATENTCODEC_LIB"] = LIBP
import numpy as np
import torch
from image_compression_2_b200 import LatentPipeline, _native

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sigma = float(sys.argv[3]) if len(sys.argv) > 3 else 0.14
lat = (torch.randn(B, 16, 512, generator=torch.Generator().manual_seed(1000 + 200000)) * sigma).cuda()
pipe = LatentPipeline(n_symbols=1 << bits)
lib = _native.load()
lib.lc_debug_profile.argtypes = [ctypes.c_void_p]
buf = np.zeros(64, np.uint64)
out = pipe.roundtrip_device(lat)
torch.cuda.synchronize()
lib.lc_debug_profile(buf.ctypes.data)  # discard the warm-up pass
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
enc = out["enc"]
ev0.record()
pipe.decode(enc.data, enc.offsets, enc.nbits, B)
ev1.record()
torch.cuda.synchronize()
assert lib.lc_debug_profile(buf.ctypes.data) == 0
t = buf.reshape(8, 8).astype(np.float64)
nsym = B * 16 * 512
print("decode launch (instrumented): %.3f ms; %d symbols" % (ev0.elapsed_time(ev1), nsym))
v2 = os.environ.get("LC_DECODER", "v2")[0] != "f" and bits <= 8
cols = ["state+wait", "symbol", "key+renorm+out", "post", "-", "-"] if v2 else ["probe", "load", "search", "interval", "renorm+out", "writeback"]
print("%-8s %8s %8s | " % ("state", "share", "cyc/sym") + " ".join("%10s" % c for c in cols))
tot = 0.0
for st in range(4):
    cnt = t[st, 0]
    if cnt == 0:
        continue
    cyc = t[st, 1:7]
    tot += cyc.sum()
    print("%-8d %7.1f%% %8.0f | " % (st, 100 * cnt / nsym, cyc.sum() / cnt) + " ".join("%10.0f" % (c / cnt) for c in cyc))
print("mean cycles/symbol/warp: %.0f" % (tot / nsym))
for st in range(4):
    cnt = max(t[st, 0], 1)
    print("state %d: exact-search fallbacks %.2f%%, exact_at fallbacks %.2f%%, waits for a recent job %.2f%%" % (
        st, 100 * t[4, st] / cnt, 100 * t[5, st] / cnt, 100 * t[6, st] / cnt))
if v2:
    busy, jobs = t[7, 0:4], t[7, 4:8]
    print("updater warp: busy cycles/job by state " + " ".join("%d: %.0f (%.1f%% of symbols)" % (i, busy[i] / max(jobs[i], 1), 100 * jobs[i] / nsym) for i in range(1, 4)))
    print("updater warp: busy cycles/symbol %.0f" % (busy.sum() / nsym))
