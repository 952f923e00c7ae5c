#!/usr/bin/env python
"""K1 / K2 / dequantisers standalone at cfg4 size against the measured HBM peak (bench.py's hbm_kernels block alone)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    r = bench.hbm_kernels(dev, peak, flush)
    for k, v in r["kernels"].items():
        print("%-82s %.3f ms  %6.1f GB/s  %.3f" % (k, v["ms"], v["achieved_gbs"], v["frac_of_measured_hbm_peak"]))
