#!/usr/bin/env python
"""Randomised parity sweep on a GPU box: CUDA encoder/decoder against the C oracle (oracle/latent_oracle.c) over random
alphabets, shapes, symbol distributions and both coder modes -- many more cases than tests/ runs each round.

  python tools/stress_parity.py [seconds] [seed]

Every case: B streams of one shape -> cabac_encode_batch (bitstreams, nbits, status, fault index equal the oracle's),
cabac_decode_batch of the ORACLE's streams (symbols up to the oracle decoder's fault, status, fault index), and the
decode of a corrupted copy.  Prints one line per failure and a summary; exit code 1 if anything differed."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image_compression_2_b200 import coder  # noqa: E402
from oracle import oracle as O  # noqa: E402


def draw_codes(rng, n, shape):
    kind = rng.integers(0, 7)
    if kind == 0:
        return rng.integers(0, n, shape)
    if kind == 1:
        sd = max(0.3, n * float(rng.choice([0.01, 0.03, 0.07, 0.15])))
        return np.clip(np.round(rng.normal(n / 2, sd, shape)), 0, n - 1)
    if kind == 2:  # two symbols alternating with noise
        a = np.zeros(shape, np.int64)
        a[..., 0::2] = rng.integers(0, n)
        a[..., 1::2] = rng.integers(0, n, a[..., 1::2].shape)
        return a
    if kind == 3:  # constant rows
        return np.broadcast_to(rng.integers(0, n, shape[:-1] + (1,)), shape).copy()
    if kind == 4:  # heavy skew
        return np.minimum((rng.exponential(max(1.0, n / 16), shape)).astype(np.int64), n - 1)
    if kind == 5:  # slowly drifting ramp
        t = np.arange(int(np.prod(shape))).reshape(shape)
        return (t // int(rng.integers(1, 40)) + rng.integers(0, 3, shape)) % n
    return np.clip(np.round(rng.normal(n * rng.random(), max(1, n / 6), shape)), 0, n - 1)


def one_case(rng, fails):
    n = int(rng.choice([2, 4, 8, 16, 32, 64, 128, 256, 256, 256, 512, 1024]))
    B = int(rng.integers(1, 6))
    R = int(rng.choice([1, 2, 3, 4, 7, 16]))
    C = int(rng.choice([1, 2, 3, 4, 5, 8, 31, 64, 100, 257, 512]))
    while R * C > 8192 or R * C * B > 20000:
        C = max(1, C // 2)
    mode = "repaired" if rng.random() < 0.8 else "verbatim"
    shape = (B, R, C)
    codes = draw_codes(rng, n, shape).astype(np.int32)
    tag = "n=%d shape=%s mode=%s" % (n, shape, mode)
    streams, nbits, status, fault = coder.cabac_encode_batch(codes, n_symbols=n, mode=mode)
    refs = [O.encode_stream(codes[b:b + 1], n, mode) for b in range(B)]
    for b, ref in enumerate(refs):
        if ref["status"] != status[b] or (ref["status"] and ref["fault_index"] != fault[b]):
            fails.append("%s b=%d encode status %d/%d fault %d/%d" % (tag, b, status[b], ref["status"], fault[b], ref["fault_index"]))
        elif ref["status"] == 0 and (nbits[b] != ref["nbits"] or streams[b] != ref["packed"]):
            fails.append("%s b=%d bitstream differs (nbits %d/%d)" % (tag, b, nbits[b], ref["nbits"]))
    good = [r["packed"] if r["status"] == 0 else b"\x00" * 8 for r in refs]
    for variant in (0, 1):
        src = list(good)
        if variant == 1:  # corrupt one byte of every stream
            src = []
            for s in good:
                s = bytearray(s)
                s[int(rng.integers(0, len(s)))] ^= int(rng.integers(1, 256))
                src.append(bytes(s))
        dec, dst, dfi = coder.cabac_decode_batch(src, shape, n_symbols=n, mode=mode)
        for b in range(B):
            rd = O.decode_stream(src[b], n, (1, R, C), mode)
            k = int(rd["fault_index"]) if rd["status"] else R * C
            if dst[b] != rd["status"] or (rd["status"] and dfi[b] != k):
                fails.append("%s b=%d v=%d decode status %d/%d fault %d/%d" % (tag, b, variant, dst[b], rd["status"], dfi[b], k))
            elif not np.array_equal(dec[b].ravel()[:k], rd["symbols"].ravel()[:k]):
                fails.append("%s b=%d v=%d decoded symbols differ" % (tag, b, variant))
    return B * R * C


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
    rng = np.random.default_rng(seed)
    fails, cases, syms, t0 = [], 0, 0, time.time()
    while time.time() - t0 < budget:
        nf = len(fails)
        syms += one_case(rng, fails)
        cases += 1
        for f in fails[nf:]:
            print("FAIL", f, flush=True)
        if len(fails) > 50:
            break
    print("stress_parity: %d cases, %d symbols, %d failures, seed %d, %.0f s" % (cases, syms, len(fails), seed, time.time() - t0))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
