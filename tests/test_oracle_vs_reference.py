"""The C oracle against the LIVE reference (only where /root/reference exists, i.e. the build
container).  Fresh random inputs each run complement the committed golden vectors."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_loader as R
from tests.helpers import ERR_TO_STATUS

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not present")


def _compare(codes, n, mode):
    bits, err = R.ref_encode_bits(codes, n, mode)
    enc = O.encode_stream(codes, n, mode)
    if err is not None:
        assert enc["status"] == ERR_TO_STATUS[err[0]] and enc["fault_index"] == err[1]
        return
    assert enc["status"] == O.OK and np.array_equal(bits, enc["bits"])
    packed = R.pack_bits(bits)
    dec, derr = R.ref_decode(packed, n, codes.shape, mode)
    mine = O.decode_stream(packed, n, codes.shape, mode)
    if derr is not None:
        assert mine["status"] == ERR_TO_STATUS[derr[0]] and mine["fault_index"] == derr[1]
    elif mine["status"] == O.DEC_NEG_SYMBOL:
        assert dec.ravel()[mine["fault_index"]] == -1
    else:
        assert mine["status"] == O.OK and np.array_equal(dec, mine["symbols"])


def test_random_small_streams_both_modes():
    rng = np.random.default_rng()
    for _ in range(60):
        n = int(rng.choice([2, 4, 16, 64, 256, 1024]))
        shape = (int(rng.integers(1, 3)), int(rng.integers(1, 5)), int(rng.integers(1, 40)))
        codes = rng.integers(0, n, shape).astype(np.int32)
        if rng.random() < 0.5:
            codes = np.clip(np.round(rng.normal(n / 2, max(1, n / 20), shape)), 0, n - 1).astype(np.int32)
        for mode in ("repaired", "verbatim"):
            _compare(codes, n, mode)


def test_one_full_stream_8bit():
    import torch
    z = (torch.randn(1, 16, 512) * 0.14).numpy()
    codes = O.quantize_codebook(z, torch.linspace(-1, 1, 256).numpy())
    _compare(codes, 256, "repaired")
