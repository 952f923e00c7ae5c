#!/usr/bin/env python
"""Encode time (quantised indices -> compacted streams) of the library named by LATENTCODEC_LIB, CUDA events, median."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_compression_2_b200 import LatentPipeline
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in (1024, 8192):
    lat = (torch.randn(B, 16, 512, generator=torch.Generator().manual_seed(1000 + 200000)) * 0.14).cuda()
    pipe = LatentPipeline(n_symbols=256)
    idx = pipe.quantize(lat)
    ts = []
    for _ in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pipe.encode(idx, reuse_output=True); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("B=%d encode %.3f ms" % (B, sorted(ts)[len(ts) // 2]), end="; ")
print()
