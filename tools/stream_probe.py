#!/usr/bin/env python
"""End-to-end throughput of LatentPipeline.roundtrip_host_stream at several depths (batches in flight) and batch sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_compression_2_b200 import LatentPipeline, codec
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for B in (1024,):
    lat = (torch.randn(B, 16, 512, generator=torch.Generator().manual_seed(5)) * 0.14).pin_memory()
    for depth, flags, eflags in ((1, 0, 0), (2, 0, 0), (3, 0, 0), (1, 0, 128), (2, 0, 128), (3, 0, 128), (2, 2, 0), (3, 2, 0)):
        codec.DEFAULT_ENCODE_FLAGS = eflags
        pipe = LatentPipeline(n_symbols=256)
        for _ in pipe.roundtrip_host_stream([lat] * (2 * depth + 2), depth=depth, dec_flags=flags):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for r in pipe.roundtrip_host_stream([lat] * steps, depth=depth, dec_flags=flags):
            pass
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        print("B=%d depth=%d dec_flags=%d enc_flags=%d  %.3f ms/batch  %.1f M symbols/s" % (B, depth, flags, eflags, 1e3 * dt, B * 8192 / dt / 1e6), flush=True)
        del pipe
        torch.cuda.empty_cache()
