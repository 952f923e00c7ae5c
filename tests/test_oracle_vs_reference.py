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
    _same_decode(dec, derr, mine)


def _same_decode(dec, derr, mine):
    """Reference result (symbols, error) against the oracle's: same symbols, same fault class, same fault index."""
    if derr is not None:
        assert mine["status"] == ERR_TO_STATUS[derr[0]] and mine["fault_index"] == derr[1], (derr, mine["status"], mine["fault_index"])
        k = derr[1]
        assert np.array_equal(np.asarray(dec).ravel()[:k], mine["symbols"].ravel()[:k])
    else:
        assert mine["status"] == O.OK and np.array_equal(dec, mine["symbols"])


def _corrupt_case(rng, trial):
    n = int(rng.choice([2, 4, 8, 16, 64, 256]))
    shape = (1, int(rng.choice([1, 2, 4])), int(rng.choice([8, 16, 48])))
    if trial % 2:
        codes = rng.integers(0, n, shape).astype(np.int32)
    else:
        codes = np.clip(np.round(rng.normal(n / 2, max(1, n / 16), shape)), 0, n - 1).astype(np.int32)
    mode = "verbatim" if trial % 3 else "repaired"
    packed = bytearray(O.encode_stream(codes, n, "repaired")["packed"])
    for _ in range(int(rng.integers(0, 4))):
        packed[int(rng.integers(0, len(packed)))] = int(rng.integers(0, 256))
    if trial % 7 == 0:
        packed = packed[: max(1, len(packed) // 2)]
    return n, shape, mode, bytes(packed)


def test_corrupt_streams_including_symbol_minus_one():
    """Corrupted streams drive the decoder through its hazards.  One of them is symbol -1 (scaled value <= 0): the
    reference does not fault there, it indexes cumulative_probs[-1] / [0], stores -1, updates probs[-1] and every
    other element, and carries on (cabac_compression.py:288-292,403) -- the oracle follows it step by step.  The
    state is nearly unreachable (a consistent decoder never has code < low; it takes the 33-bit values of the
    verbatim mode), so the oracle first searches many random corruptions for streams that pass through it, and the
    live reference is then run on those and on a sample of the others."""
    rng = np.random.default_rng(3)
    hits, others = [], []
    for trial in range(40000):
        n, shape, mode, packed = _corrupt_case(rng, trial)
        mine = O.decode_stream(packed, n, shape, mode)
        k = mine["fault_index"] if mine["status"] else int(np.prod(shape))
        if (mine["symbols"].ravel()[:k] == -1).any():
            hits.append((n, shape, mode, packed))
        elif trial % 100 == 0:
            others.append((n, shape, mode, packed))
    assert len(hits) >= 3, len(hits)
    for n, shape, mode, packed in hits[:12] + others:
        dec, derr = R.ref_decode(packed, n, shape, mode)
        mine = O.decode_stream(packed, n, shape, mode)
        _same_decode(dec, derr, mine)


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


def _ref_model_dict(cm):
    """The reference object's state as {(left, up): (vector, count)}; `()` (non-3-D data) maps to (-2, -2)."""
    out = {}
    for k, v in cm.context_models.items():
        kk = (-2, -2) if len(k) == 0 else (int(k[0]), int(k[1]))
        out[kk] = (np.asarray(v, np.float64), int(cm.context_counts.get(k, 0)))
    return out


def _same_model(a, b):
    assert set(a) == set(b)
    for k in a:
        assert a[k][1] == b[k][1], k
        assert np.array_equal(a[k][0], b[k][0]), k


def test_shared_model_across_calls_defect_d5():
    """One ContextModel mutated by successive cabac_encode / cabac_decode calls, as CABACCompressor does
    (cabac_compression.py:438,478,517): bits, decoded symbols and the model object's state after every call."""
    cc = R.set_mode("repaired")
    rng = np.random.default_rng(11)
    for n, shape in ((16, (2, 3, 30)), (64, (1, 4, 48)), (256, (1, 2, 64)), (16, (50,))):
        a = np.clip(np.round(rng.normal(n / 2, max(1, n / 16), shape)), 0, n - 1).astype(np.int32)
        b = np.clip(np.round(rng.normal(n / 2, max(1, n / 16), shape)), 0, n - 1).astype(np.int32)
        cm = cc.ContextModel(n_symbols=n)
        model = {}
        for codes in (a, b):
            bits = np.frombuffer(cc.cabac_encode(codes, cm), np.uint8)
            mine, model = O.encode_stream_model(codes, n, model)
            assert mine["status"] == O.OK and np.array_equal(bits, mine["bits"])
            _same_model(_ref_model_dict(cm), model)
        # decoding with the model the encoder left behind (what CABACCompressor.decompress does): usually garbage
        packed = O.encode_stream(a, n)["packed"]
        dec, derr = None, None
        try:
            dec = cc.cabac_decode(packed, cm, a.shape)
        except (IndexError, ZeroDivisionError, ValueError) as e:
            derr = type(e).__name__
        mine, model2 = O.decode_stream_model(packed, n, a.shape, model)
        if derr is None:
            assert mine["status"] == O.OK and np.array_equal(np.asarray(dec, np.int32), mine["symbols"])
            _same_model(_ref_model_dict(cm), model2)
        else:
            assert mine["status"] == ERR_TO_STATUS[derr]
