"""The coder DEVICE code (image_compression_2_b200/csrc/lc_coder.cuh), executed on the CPU by the
SIMT emulator in tests/hostsim, against the golden vectors of the reference and against the C
oracle.  This is how the warp-cooperative logic is checked in the GPU-less build container; the
same comparisons run on the real kernels in tests/test_gpu_*.py."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import ERR_TO_STATUS, coder_cases, golden
from tests.hostsim import build as H

MODE = {"verbatim": 0, "repaired": 1}


def _pow2(n):
    return n >= 2 and (n & (n - 1)) == 0


def _as_batch(codes):
    """golden streams are (B,R,C) coded as ONE stream (B images share the model) or 1-D."""
    if codes.ndim == 3:
        return codes[None]  # (1 stream, imgs, R, C)
    return codes.reshape(1, -1)  # single global context


def _check(rec):
    n = int(rec["n"])
    codes = rec["codes"]
    mode = MODE[rec["mode"]]
    batch = _as_batch(codes)
    out, nbits, status, fault = H.encode(batch, n, mode)
    if "enc_error" in rec:
        assert status[0] == ERR_TO_STATUS[str(rec["enc_error"][0])]
        assert fault[0] == int(rec["enc_fault_index"])
        return
    assert status[0] == 0
    assert nbits[0] == int(rec["nbits"])
    packed = rec["packed"].tobytes()
    assert out[0, : len(packed)].tobytes() == packed
    dec, dstatus, dfault, _ = H.decode([packed], n, batch.shape, mode)
    ref_dec = rec["decoded"].reshape(batch.shape)
    if "dec_error" in rec:
        k = int(rec["dec_fault_index"])
        assert dstatus[0] == ERR_TO_STATUS[str(rec["dec_error"][0])] and dfault[0] == k
        assert np.array_equal(dec.ravel()[:k], ref_dec.ravel()[:k])
        assert not dec.ravel()[k:].any()
    elif dstatus[0] == 4:
        k = dfault[0]
        assert ref_dec.ravel()[k] == -1 and np.array_equal(dec.ravel()[:k], ref_dec.ravel()[:k])
    else:
        assert dstatus[0] == 0
        assert np.array_equal(dec, ref_dec)


@pytest.mark.parametrize("fixture", ["kat.npz", "coder_small.npz", "coder_full.npz"])
def test_device_code_matches_reference_vectors(fixture):
    cases = coder_cases(golden(fixture))
    ran = 0
    for name, rec in cases.items():
        if not _pow2(int(rec["n"])):
            continue
        try:
            _check(rec)
        except AssertionError as e:
            raise AssertionError("case %s: %s" % (name, e))
        ran += 1
    assert ran > 0


def test_config1_stream():
    c = golden("config1.npz")
    rec = {k.split("__", 1)[1]: c[k] for k in c.files if k.startswith("coder_repaired__")}
    rec["mode"] = "repaired"
    _check(rec)
    rec = {k.split("__", 1)[1]: c[k] for k in c.files if k.startswith("coder_verbatim__")}
    rec["mode"] = "verbatim"
    _check(rec)


def test_batch_of_independent_streams_and_fused_dequant():
    rng = np.random.default_rng(5)
    n = 64
    codes = np.clip(np.round(rng.normal(32, 5, (5, 4, 96))), 0, n - 1).astype(np.int32)
    out, nbits, status, fault = H.encode(codes, n, 1, grid=2)
    streams = []
    for b in range(5):
        r = O.encode_stream(codes[b:b + 1], n)
        assert status[b] == 0 and nbits[b] == r["nbits"]
        assert out[b, : len(r["packed"])].tobytes() == r["packed"]
        streams.append(r["packed"])
    cb = np.linspace(-1, 1, n).astype(np.float32)
    dec, dstatus, _, deq = H.decode(streams, n, codes.shape, 1, grid=2, codebook=cb)
    assert not dstatus.any() and np.array_equal(dec, codes)
    assert np.array_equal(deq.reshape(codes.shape), cb[codes])


def test_output_slot_overflow_is_reported():
    rng = np.random.default_rng(6)
    codes = rng.integers(0, 256, (1, 4, 64)).astype(np.int32)
    out, nbits, status, fault = H.encode(codes, 256, 1, slot_bytes=64)
    assert status[0] == 5


def test_bad_symbol_is_reported():
    codes = np.zeros((1, 2, 40), np.int32)
    codes[0, 1, 3] = 16
    out, nbits, status, fault = H.encode(codes, 16, 1)
    assert status[0] == 6 and fault[0] == 43
