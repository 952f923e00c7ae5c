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


# ---- parallel encoder (phase A / phase B device code; phase S restated on the host) ---------------

def test_parallel_encoder_matches_reference_vectors():
    ran = 0
    for fixture in ("kat.npz", "coder_small.npz", "coder_full.npz"):
        for name, rec in coder_cases(golden(fixture)).items():
            n, codes = int(rec["n"]), rec["codes"]
            if not _pow2(n) or codes.ndim != 3 or codes.size > 8192:
                continue
            out, nbits, status, fault = H.encode_par(codes[None], n, MODE[rec["mode"]], nwarps=1 + ran % 4)
            if "enc_error" in rec:
                assert status[0] == ERR_TO_STATUS[str(rec["enc_error"][0])], name
                assert fault[0] == int(rec["enc_fault_index"]), name
            else:
                packed = rec["packed"].tobytes()
                assert status[0] == 0 and nbits[0] == int(rec["nbits"]), name
                assert out[0, : len(packed)].tobytes() == packed, name
            ran += 1
    assert ran > 200


def test_parallel_encoder_batches_and_faults():
    rng = np.random.default_rng(21)
    for n, shape in ((16, (4, 16, 512)), (256, (3, 16, 512)), (1024, (2, 16, 512)), (64, (5, 3, 100))):
        codes = np.clip(np.round(rng.normal(n / 2, max(1.0, n / 14), shape)), 0, n - 1).astype(np.int32)
        out, nbits, status, fault = H.encode_par(codes, n, 1, grid=2, nwarps=4)
        for b in range(shape[0]):
            ref = O.encode_stream(codes[b:b + 1], n, "repaired")
            assert status[b] == 0 and nbits[b] == ref["nbits"]
            assert out[b, : len(ref["packed"])].tobytes() == ref["packed"]
    # a bad symbol stops the stream exactly where the serial encoder stops
    codes = rng.integers(0, 16, (2, 4, 64)).astype(np.int32)
    codes[1, 2, 7] = 99
    out, nbits, status, fault = H.encode_par(codes, 16, 1)
    assert status.tolist() == [0, 6] and fault[1] == 2 * 64 + 7
    # verbatim mode: the D3 fault precedes a later bad symbol, as in the serial order
    codes = np.clip(np.round(rng.normal(128, 20, (1, 16, 512))), 0, 255).astype(np.int32)
    codes[0, 15, 500] = 300
    ref = O.encode_stream(np.where(codes > 255, 0, codes), 256, "verbatim")
    out, nbits, status, fault = H.encode_par(codes, 256, 0)
    assert ref["status"] == 1 and status[0] == 1 and fault[0] == ref["fault_index"]
    # tiny output slot
    out, nbits, status, fault = H.encode_par(rng.integers(0, 256, (1, 4, 64)).astype(np.int32), 256, 1, slot_bytes=64)
    assert status[0] == 5


# ---- fast decoder (lazy states, register-resident model, closed-form renormalisation) ----------------

def test_fast_decoder_matches_reference_vectors():
    ran = 0
    for fixture in ("kat.npz", "coder_small.npz", "coder_full.npz"):
        for name, rec in coder_cases(golden(fixture), mode="repaired").items():
            n, codes = int(rec["n"]), rec["codes"]
            if not _pow2(n) or codes.ndim != 3 or "enc_error" in rec:
                continue
            batch = codes[None]
            dec, st, fi, _ = H.decode([rec["packed"].tobytes()], n, batch.shape, 1, fast=True)
            ref = rec["decoded"].reshape(batch.shape)
            if "dec_error" in rec:
                k = int(rec["dec_fault_index"])
                assert st[0] == ERR_TO_STATUS[str(rec["dec_error"][0])] and fi[0] == k, name
                assert np.array_equal(dec.ravel()[:k], ref.ravel()[:k]) and not dec.ravel()[k:].any(), name
            else:
                assert st[0] == 0 and np.array_equal(dec, ref), name
            ran += 1
    assert ran > 100


def test_fast_decoder_hands_wide_contexts_to_generic_kernel():
    """A context with more than 32 distinct symbols: the fast kernel flags the stream, the generic kernel redoes it."""
    rng = np.random.default_rng(31)
    n = 256
    codes = np.zeros((2, 4, 400), np.int32)
    codes[0, :, 0::2] = 7                              # (left=7, up=7|-1) contexts recur ...
    codes[0, :, 1::2] = rng.integers(0, n, (4, 200))   # ... with many different symbols
    codes[1] = np.clip(np.round(rng.normal(128, 9, (4, 400))), 0, n - 1)
    streams = [O.encode_stream(codes[b:b + 1], n)["packed"] for b in range(2)]
    cb = np.linspace(-1, 1, n).astype(np.float32)
    dec, st, fi, deq = H.decode(streams, n, codes.shape, 1, fast=True, grid=1, codebook=cb)
    assert H.decode.last_redone == 1
    assert not st.any() and np.array_equal(dec, codes) and np.array_equal(deq.reshape(codes.shape), cb[codes])


def test_lane_per_group_phase_a():
    """nwarps=0 selects the lane-per-group phase A (n <= 256): same bitstreams as the oracle."""
    rng = np.random.default_rng(41)
    c = golden("config1.npz")
    out, nbits, status, _ = H.encode_par(c["coder_repaired__codes"][None], 256, 1, nwarps=0)
    packed = c["coder_repaired__packed"].tobytes()
    assert status[0] == 0 and out[0, : len(packed)].tobytes() == packed
    for n, shape in ((2, (2, 3, 40)), (4, (2, 4, 64)), (16, (3, 16, 512)), (64, (4, 5, 100)), (128, (2, 8, 200)), (256, (2, 16, 512))):
        codes = np.clip(np.round(rng.normal(n / 2, max(0.8, n / 14), shape)), 0, n - 1).astype(np.int32)
        codes[0, 0, :8] = n - 1
        out, nbits, status, fault = H.encode_par(codes, n, 1, grid=2, nwarps=0)
        for b in range(shape[0]):
            ref = O.encode_stream(codes[b:b + 1], n, "repaired")
            assert status[b] == 0 and nbits[b] == ref["nbits"] and out[b, : len(ref["packed"])].tobytes() == ref["packed"]
    codes = rng.integers(0, 16, (2, 4, 64)).astype(np.int32)
    codes[1, 2, 7] = 99
    out, nbits, status, fault = H.encode_par(codes, 16, 1, nwarps=0)
    assert status.tolist() == [0, 6] and fault[1] == 2 * 64 + 7
    # contexts that collect more than 32 distinct symbols continue on the dense image
    for n in (256, 1024):
        codes = np.zeros((2, 4, 400), np.int32)
        codes[0, :, 0::2] = 7
        codes[0, :, 1::2] = rng.integers(0, n, (4, 200))
        codes[1] = np.clip(np.round(rng.normal(n / 2, n / 28, (4, 400))), 0, n - 1)
        out, nbits, status, fault = H.encode_par(codes, n, 1, grid=2, nwarps=0)
        for b in range(2):
            ref = O.encode_stream(codes[b:b + 1], n, "repaired")
            assert status[b] == 0 and nbits[b] == ref["nbits"] and out[b, : len(ref["packed"])].tobytes() == ref["packed"]


# ---- decoder v2 (decoder warp + updater warps, direct-mapped contexts, integer / table fast paths) ----

def _v2_ok(n, shape):
    return 8 <= n <= 256 and len(shape) == 3 and 4 <= shape[2] <= 4096


@pytest.mark.parametrize("ver", ["v2", "v3"])
def test_decoder_v2_matches_reference_vectors(ver):
    ran = 0
    for fixture in ("kat.npz", "coder_small.npz", "coder_full.npz"):
        for name, rec in coder_cases(golden(fixture), mode="repaired").items():
            n, codes = int(rec["n"]), rec["codes"]
            if not _pow2(n) or codes.ndim != 3 or "enc_error" in rec or not _v2_ok(n, codes.shape):
                continue
            batch = codes[None]
            dec, st, fi, _ = H.decode([rec["packed"].tobytes()], n, batch.shape, 1, fast=ver)
            ref = rec["decoded"].reshape(batch.shape)
            if "dec_error" in rec:
                k = int(rec["dec_fault_index"])
                assert st[0] == ERR_TO_STATUS[str(rec["dec_error"][0])] and fi[0] == k, name
                assert np.array_equal(dec.ravel()[:k], ref.ravel()[:k]) and not dec.ravel()[k:].any(), name
            else:
                assert st[0] == 0 and np.array_equal(dec, ref), name
            ran += 1
    assert ran > 40


@pytest.mark.parametrize("ver", ["v2", "v3"])
def test_decoder_v2_against_oracle_decoder(ver):
    """Random streams over the alphabets v2 accepts, including narrow 4-bit data the reference itself cannot
    round-trip (hazards H1-H3): same symbols, same fault class, same fault index as the oracle decoder."""
    rng = np.random.default_rng(77)
    cases = ((256, (3, 16, 512), 18.0), (256, (2, 16, 512), 2.5), (128, (2, 8, 200), 9.0), (64, (4, 5, 100), 4.5),
             (16, (3, 16, 512), 1.14), (16, (2, 16, 512), 0.6), (8, (3, 6, 64), 1.0), (256, (2, 2, 2000), 30.0))
    for n, shape, sd in cases:
        codes = np.clip(np.round(rng.normal(n / 2, sd, shape)), 0, n - 1).astype(np.int32)
        codes[0, 0, :8] = n - 1
        streams = [O.encode_stream(codes[b:b + 1], n)["packed"] for b in range(shape[0])]
        cb = np.linspace(-1, 1, n).astype(np.float32)
        dec, st, fi, deq = H.decode(streams, n, codes.shape, 1, fast=ver, grid=2, codebook=cb)
        for b in range(shape[0]):
            ref = O.decode_stream(streams[b], n, (1,) + shape[1:])
            k = int(ref["fault_index"]) if ref["status"] else codes[b].size
            assert st[b] == ref["status"] and (ref["status"] == 0 or fi[b] == k), (n, shape, b, st[b], ref["status"])
            assert np.array_equal(dec[b].ravel()[:k], ref["symbols"].ravel()[:k]), (n, shape, b)
            assert np.array_equal(deq[b][:k], cb[dec[b].ravel()[:k]])


@pytest.mark.parametrize("ver", ["v2", "v3"])
def test_decoder_v2_wide_contexts_and_images_sharing_a_model(ver):
    rng = np.random.default_rng(31)
    n = 256
    codes = np.zeros((2, 4, 400), np.int32)
    codes[0, :, 0::2] = 7                              # (left=7, up=7|-1) contexts recur ...
    codes[0, :, 1::2] = rng.integers(0, n, (4, 200))   # ... with many different symbols: > 32 entries
    codes[1] = np.clip(np.round(rng.normal(128, 9, (4, 400))), 0, n - 1)
    streams = [O.encode_stream(codes[b:b + 1], n)["packed"] for b in range(2)]
    cb = np.linspace(-1, 1, n).astype(np.float32)
    dec, st, fi, deq = H.decode(streams, n, codes.shape, 1, fast=ver, grid=1, codebook=cb)
    assert H.decode.last_redone == 1
    assert not st.any() and np.array_equal(dec, codes) and np.array_equal(deq.reshape(codes.shape), cb[codes])
    # records between 7 and 32 entries live in the pool
    codes = np.zeros((1, 6, 300), np.int32)
    codes[0, :, 0::2] = 5
    codes[0, :, 1::2] = rng.integers(100, 120, (6, 150))
    streams = [O.encode_stream(codes, n)["packed"]]
    dec, st, fi, _ = H.decode(streams, n, codes.shape, 1, fast=ver, grid=1)
    assert H.decode.last_redone == 0 and not st.any() and np.array_equal(dec, codes)
    # the reference's batched call: several images share coder and model (cabac_compression.py:330-337)
    imgs = np.clip(np.round(rng.normal(32, 3, (3, 4, 64))), 0, 63).astype(np.int32)
    packed = O.encode_stream(imgs, 64)["packed"]
    dec, st, fi, _ = H.decode([packed], 64, (1,) + imgs.shape, 1, fast=ver, grid=1)
    assert not st.any() and np.array_equal(dec[0], imgs)
    # corrupted stream: whatever the oracle decoder does, v2 does
    bad = bytearray(packed); bad[len(bad) // 2] ^= 0x5a
    ref = O.decode_stream(bytes(bad), 64, imgs.shape)
    dec, st, fi, _ = H.decode([bytes(bad)], 64, (1,) + imgs.shape, 1, fast=ver, grid=1)
    k = int(ref["fault_index"]) if ref["status"] else imgs.size
    assert st[0] == ref["status"] and np.array_equal(dec.ravel()[:k], ref["symbols"].ravel()[:k])


def test_decoder_v2_truncated_streams_read_zeros():
    """Valid streams cut at every byte alignment (the flat loop's bit reader keeps its next word pre-masked and loads a
    word that straddles the end of the stream on its rare path): symbols, status and fault index of the oracle decoder."""
    rng = np.random.default_rng(5)
    n, shape = 256, (1, 4, 128)
    codes = np.clip(np.round(rng.normal(128, 18, shape)), 0, n - 1).astype(np.int32)
    full = bytes(O.encode_stream(codes, n)["packed"])
    cuts = sorted(set(list(range(0, 10)) + [len(full) - k for k in range(0, 7)] + [len(full) // 2 + k for k in range(4)]))
    streams = [full[:c] for c in cuts]
    dec, st, fi, _ = H.decode(streams, n, (len(streams),) + shape[1:], 1, fast="v2", grid=2)
    for b, s in enumerate(streams):
        ref = O.decode_stream(s, n, shape)
        k = int(ref["fault_index"]) if ref["status"] else codes.size
        assert st[b] == ref["status"] and (ref["status"] == 0 or fi[b] == k), (len(s), st[b], ref["status"])
        assert np.array_equal(dec[b].ravel()[:k], ref["symbols"].ravel()[:k]), len(s)


# ---- small-alphabet decoder (dense shared-memory model, n <= 16) ----

def test_small_decoder_matches_reference_vectors():
    ran = 0
    for fixture in ("kat.npz", "coder_small.npz", "coder_full.npz"):
        for name, rec in coder_cases(golden(fixture), mode="repaired").items():
            n, codes = int(rec["n"]), rec["codes"]
            if not _pow2(n) or n > 16 or codes.ndim != 3 or "enc_error" in rec:
                continue
            batch = codes[None]
            dec, st, fi, _ = H.decode([rec["packed"].tobytes()], n, batch.shape, 1, fast="small")
            ref = rec["decoded"].reshape(batch.shape)
            if "dec_error" in rec:
                k = int(rec["dec_fault_index"])
                assert st[0] == ERR_TO_STATUS[str(rec["dec_error"][0])] and fi[0] == k, name
                assert np.array_equal(dec.ravel()[:k], ref.ravel()[:k]) and not dec.ravel()[k:].any(), name
            else:
                assert st[0] == 0 and np.array_equal(dec, ref), name
            ran += 1
    assert ran > 10, ran


def test_small_decoder_against_oracle_decoder():
    """Random streams at 1..4 bits, including narrow 4-bit data the reference cannot round-trip (hazards H1-H3), several
    images sharing a model, short rows and a corrupted stream: same symbols, fault class and fault index as the oracle."""
    rng = np.random.default_rng(1234)
    cases = ((16, (3, 16, 512), 1.14), (16, (2, 16, 512), 0.6), (16, (2, 16, 512), 3.0), (8, (3, 6, 64), 1.0),
             (4, (2, 5, 33), 0.8), (2, (2, 3, 40), 0.5), (16, (4, 1, 7), 2.0), (8, (2, 9, 1), 1.5))
    for n, shape, sd in cases:
        codes = np.clip(np.round(rng.normal(n / 2, sd, shape)), 0, n - 1).astype(np.int32)
        codes[0, 0, :3] = n - 1
        streams = [O.encode_stream(codes[b:b + 1], n)["packed"] for b in range(shape[0])]
        bad = bytearray(streams[-1])
        if len(bad) > 4:
            bad[len(bad) // 2] ^= 0x5a
            streams[-1] = bytes(bad)
        cb = np.linspace(-1, 1, n).astype(np.float32)
        dec, st, fi, deq = H.decode(streams, n, codes.shape, 1, fast="small", grid=2, codebook=cb)
        for b in range(shape[0]):
            ref = O.decode_stream(streams[b], n, (1,) + shape[1:])
            k = int(ref["fault_index"]) if ref["status"] else codes[b].size
            assert st[b] == ref["status"] and (ref["status"] == 0 or fi[b] == k), (n, shape, b, st[b], ref["status"])
            assert np.array_equal(dec[b].ravel()[:k], ref["symbols"].ravel()[:k]), (n, shape, b)
            assert np.array_equal(deq[b][:k], cb[dec[b].ravel()[:k]])
    imgs = np.clip(np.round(rng.normal(8, 2, (3, 4, 64))), 0, 15).astype(np.int32)
    packed = O.encode_stream(imgs, 16)["packed"]
    dec, st, fi, _ = H.decode([packed], 16, (1,) + imgs.shape, 1, fast="small", grid=1)
    assert not st.any() and np.array_equal(dec[0], imgs)


def test_symbol_minus_one_is_followed_like_the_reference():
    """A decoded symbol -1 is not a fault in the reference (NumPy negative indexing, cabac_compression.py:288-292,403).
    The state is nearly unreachable, so the oracle (pinned against the live reference on exactly this in
    tests/test_oracle_vs_reference.py) searches random corruptions for streams that pass through it; the device code
    -- generic kernel, and the fast kernels handing over to it -- must give the oracle's symbols, status, fault index."""
    rng = np.random.default_rng(3)
    hits = []
    for trial in range(40000):
        n = int(rng.choice([2, 4, 8, 16, 64, 256]))
        shape = (1, int(rng.choice([1, 2, 4])), int(rng.choice([8, 16, 48])))
        codes = rng.integers(0, n, shape).astype(np.int32)
        mode = "verbatim" if trial % 3 else "repaired"
        packed = bytearray(O.encode_stream(codes, n, "repaired")["packed"])
        for _ in range(int(rng.integers(0, 4))):
            packed[int(rng.integers(0, len(packed)))] = int(rng.integers(0, 256))
        ref = O.decode_stream(bytes(packed), n, shape, mode)
        k = ref["fault_index"] if ref["status"] else codes.size
        if (ref["symbols"].ravel()[:k] == -1).any():
            hits.append((n, shape, mode, bytes(packed), ref, k))
    assert len(hits) >= 3
    for n, shape, mode, packed, ref, k in hits[:10]:
        cb = np.linspace(-1, 1, n).astype(np.float32)
        dec, st, fi, deq = H.decode([packed], n, (1,) + shape, 0 if mode == "verbatim" else 1, codebook=cb)
        assert st[0] == ref["status"] and (ref["status"] == 0 or fi[0] == k), (n, shape, mode)
        assert np.array_equal(dec.ravel()[:k], ref["symbols"].ravel()[:k]), (n, shape, mode)
        assert np.array_equal(deq.ravel()[:k], cb[ref["symbols"].ravel()[:k]])  # codebook[-1] is the last entry
        if mode == "repaired" and shape[2] >= 4:
            for fast in (True, "v2", "small"):
                if (fast == "small" and n > 16) or (fast == "v2" and not _v2_ok(n, (1,) + shape[1:])):
                    continue
                dec, st, fi, _ = H.decode([packed], n, (1,) + shape, 1, fast=fast)
                assert st[0] == ref["status"] and np.array_equal(dec.ravel()[:k], ref["symbols"].ravel()[:k]), (n, shape, fast)
