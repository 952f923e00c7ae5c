"""Shared helpers for the test-suite (golden fixture access, synthetic latents)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ERR_TO_STATUS = {"ValueError": 1, "IndexError": 2, "ZeroDivisionError": 3}


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def coder_cases(npz, mode=None):
    """Group 'case__field' keys of a coder fixture into {case: {field: array}}."""
    cases = {}
    for k in npz.files:
        if "__" not in k or k.startswith("_"):
            continue
        case, field = k.split("__", 1)
        cases.setdefault(case, {})[field] = npz[k]
    out = {}
    for case, rec in cases.items():
        if "codes" not in rec:
            continue
        m = "verbatim" if case.endswith("_verbatim") else "repaired"
        if mode is not None and m != mode:
            continue
        rec["mode"] = m
        out[case] = rec
    return out


def synth_latents(kind, B, seed, R=16, C=512):
    """SURVEY.md section 8d synthetic W+ latents, generated on the CPU so every path sees the same bits."""
    g = torch.Generator().manual_seed(seed)
    if kind == "enc_like":
        return torch.randn(B, R, C, generator=g) * 0.14
    if kind == "wide":
        return torch.randn(B, R, C, generator=g) * 0.4
    if kind == "uniform":
        return torch.rand(B, R, C, generator=g) * 2 - 1
    if kind == "hier":
        sig = torch.empty(R)
        sig[:5], sig[5:12], sig[12:] = 0.115, 0.158, 0.131
        return torch.randn(B, R, C, generator=g) * sig.view(1, R, 1)
    raise ValueError(kind)
