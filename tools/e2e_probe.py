#!/usr/bin/env python
"""End-to-end (host buffers in and out) time of LatentPipeline.roundtrip_host by chunk count, beside the device-resident
step.  python tools/e2e_probe.py [B ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from image_compression_2_b200 import LatentPipeline  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in [int(a) for a in sys.argv[1:]] or [1024, 8192]:
    lat_host = (torch.randn(B, 16, 512, generator=torch.Generator().manual_seed(5)) * 0.14).pin_memory()
    lat = lat_host.cuda()
    pipe = LatentPipeline(n_symbols=256)
    for _ in range(3):
        pipe.roundtrip_device(lat)
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.roundtrip_device(lat)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("B=%d device step %.3f ms" % (B, sorted(ts)[len(ts) // 2]))
    for chunks in (1, 2, 3, 4):
        for _ in range(3):
            pipe.roundtrip_host(lat_host, chunks=chunks)
        ts = []
        for _ in range(10):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipe.roundtrip_host(lat_host, chunks=chunks)
            ts.append(1e3 * (time.perf_counter() - t0))
        print("B=%d chunks=%d e2e %.3f ms (min %.3f)" % (B, chunks, sorted(ts)[len(ts) // 2], min(ts)))
