"""The C oracle (oracle/latent_oracle.c) against the golden vectors produced by executing the
reference itself (oracle/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import ERR_TO_STATUS, coder_cases, golden


def _check_case(rec):
    mode = rec["mode"]
    n = int(rec["n"])
    codes = rec["codes"]
    enc = O.encode_stream(codes, n, mode)
    if "enc_error" in rec:
        assert enc["status"] == ERR_TO_STATUS[str(rec["enc_error"][0])]
        assert enc["fault_index"] == int(rec["enc_fault_index"])
        return
    assert enc["status"] == O.OK
    assert enc["nbits"] == int(rec["nbits"])
    assert enc["packed"] == rec["packed"].tobytes()
    dec = O.decode_stream(rec["packed"].tobytes(), n, codes.shape, mode)
    if "dec_error" in rec:
        assert dec["status"] == ERR_TO_STATUS[str(rec["dec_error"][0])]
        k = int(rec["dec_fault_index"])
        assert dec["fault_index"] == k
        assert np.array_equal(dec["symbols"].ravel()[:k], rec["decoded"].ravel()[:k])
    else:
        # (streams whose decoding passes through symbol -1 carry on with NumPy's negative indexing, like the reference)
        assert dec["status"] == O.OK
        assert np.array_equal(dec["symbols"], rec["decoded"])


@pytest.mark.parametrize("fixture", ["kat.npz", "coder_full.npz", "coder_small.npz"])
def test_coder_matches_reference_vectors(fixture):
    cases = coder_cases(golden(fixture))
    assert cases
    for name, rec in cases.items():
        try:
            _check_case(rec)
        except AssertionError as e:
            raise AssertionError("case %s: %s" % (name, e))


def test_survey_known_answers():
    k = golden("kat.npz")
    bits = np.unpackbits(k["kat1_verbatim__packed"])[: int(k["kat1_verbatim__nbits"])]
    assert "".join(map(str, bits)) == "0011100100111001000001100101"
    bits = np.unpackbits(k["kat1_repaired__packed"])[: int(k["kat1_repaired__nbits"])]
    assert "".join(map(str, bits)) == "00111001001101111110111110110010"
    for m in ("verbatim", "repaired"):
        assert int(k["kat2_%s__nbits" % m]) == 258
        assert hashlib.sha256(k["kat2_%s__packed" % m].tobytes()).hexdigest()[:16] == "13293a63328aed76"
    assert int(k["kat3_repaired__nbits"]) == 65600
    c = golden("config1.npz")
    assert int(c["coder_repaired__nbits"]) == 65697
    assert int(c["coder_verbatim__enc_fault_index"]) == 294


def test_config1_quantisers_and_coder():
    c = golden("config1.npz")
    means = c["means"]
    for bits in (4, 8, 10):
        idx, wq = O.quantize_affine(means, bits)
        assert np.array_equal(idx, c["a_idx_%d" % bits].astype(np.int32))
        assert np.array_equal(wq.view(np.uint32), c["a_wq_%d" % bits].view(np.uint32))
        assert np.array_equal(O.dequantize_affine(idx, bits).view(np.uint32), c["a_wq_%d" % bits].view(np.uint32))
    idx = O.quantize_codebook(means, c["codebook_256"])
    assert np.array_equal(idx, c["b_idx_256"])
    assert np.array_equal(O.dequantize_codebook(idx, c["codebook_256"]), c["b_deq_256"])
    enc = O.encode_stream(idx, 256, "repaired")
    assert enc["packed"] == c["coder_repaired__packed"].tobytes() and enc["n_contexts"] == 3900


def test_quantiser_fixtures():
    q = golden("quantizers.npz")
    w = q["w"]
    for bits in (4, 6, 8, 10):
        idx, wq = O.quantize_affine(w, bits)
        ref_idx = q["a_idx_%d" % bits]
        fin = np.isfinite(ref_idx) & (np.abs(ref_idx) < 2e9)
        assert np.array_equal(idx[fin], ref_idx[fin].astype(np.int32))
        assert np.array_equal(wq.view(np.uint32)[fin], q["a_wq_%d" % bits].view(np.uint32)[fin])
        assert np.array_equal(np.isnan(wq), np.isnan(q["a_wq_%d" % bits]))
    for n in (16, 64, 256, 1024):
        idx = O.quantize_codebook(q["b_z_%d" % n], q["codebook_%d" % n])
        assert np.array_equal(idx, q["b_idx_%d" % n])
        assert np.array_equal(O.dequantize_codebook(idx, q["codebook_%d" % n]), q["b_deq_%d" % n])


def test_numpy_pairwise_sum_model():
    rng = np.random.default_rng(3)
    for n in (2, 4, 7, 8, 16, 100, 128, 129, 256, 512, 1000, 1024):
        for _ in range(100):
            a = rng.random(n) ** 3
            assert O.np_sum(a) == a.sum()
