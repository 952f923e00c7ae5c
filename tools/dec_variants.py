#!/usr/bin/env python
"""Decode time of liblatentcodec builds that differ in compile-time switches of the decoder (-D...), on one GPU box.

  python tools/dec_variants.py build   (here: cross-compiles tools/variants/*.so)
  python tools/dec_variants.py run     (GPU box: one subprocess per library, prints decode / step ms at 1024 and 8192 streams)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "tools", "variants")
VARIANTS = {
    "old": ("LCV_OPT_FLAT=0",),
    "flat": (),
    "flat_outline": ("LCV_OPT_OUTLINE_LAT=true",),
    "flat_all": ("LCVF_NO_SYNCWARP=1", "LCV_OPT_ROLE_SWAP=1"),
}

if len(sys.argv) > 1 and sys.argv[1] == "build":
    from image_compression_2_b200 import build
    os.makedirs(VDIR, exist_ok=True)
    extra = tuple(sys.argv[2:])
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(4) as ex:
        for r in ex.map(lambda kv: build.build_library(out=os.path.join(VDIR, "lib_%s.so" % kv[0]), defines=kv[1] + extra),
                        VARIANTS.items()):
            print(r)
    sys.exit(0)

if len(sys.argv) > 1 and sys.argv[1] == "run":
    for name in sorted(os.listdir(VDIR)):
        if name.endswith(".so"):
            env = dict(os.environ, LATENTCODEC_LIB=os.path.join(VDIR, name))
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, capture_output=True, text=True)
            print("%-22s %s" % (name, out.stdout.strip() or out.stderr.strip()[-300:]), flush=True)
    sys.exit(0)

import torch  # noqa: E402

from image_compression_2_b200 import LatentPipeline  # noqa: E402

res = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in (1024, 8192):
    lat = (torch.randn(B, 16, 512, generator=torch.Generator().manual_seed(1000 + 200000)) * 0.14).cuda()
    pipe = LatentPipeline(n_symbols=256)
    out = pipe.roundtrip_device(lat)
    torch.cuda.synchronize()
    assert int(out["dec_status"].abs().sum()) == 0 and torch.equal(out["dec_idx"], out["idx"]), "round trip broken"
    enc = out["enc"]
    td, ts = [], []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.decode(enc.data, enc.offsets, enc.nbits, B)
        e1.record()
        e1.synchronize()
        td.append(e0.elapsed_time(e1))
    res.append("B=%d decode %.3f ms" % (B, sorted(td)[len(td) // 2]))
print("; ".join(res))
