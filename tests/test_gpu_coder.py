"""K3/K4/K5 on the B200 through the C ABI and the drop-in functions: bit-exact against the golden
vectors the reference produced, and against the C oracle on seeded synthetic latents."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.helpers import ERR_TO_STATUS, coder_cases, golden, synth_latents

pytestmark = pytest.mark.gpu

EXC = {"ValueError": ValueError, "IndexError": IndexError, "ZeroDivisionError": ZeroDivisionError}


def _pow2(n):
    return n >= 2 and (n & (n - 1)) == 0


def _check_case(rec):
    from image_compression_2_b200 import coder
    n, codes, mode = int(rec["n"]), rec["codes"], rec["mode"]
    cm = coder.ContextModel(n_symbols=n)
    if "enc_error" in rec:
        with pytest.raises(EXC[str(rec["enc_error"][0])]) as ei:
            coder.cabac_encode(codes, cm, mode=mode)
        assert "at symbol %d" % int(rec["enc_fault_index"]) in str(ei.value)
        return
    bits = coder.cabac_encode(codes, cm, mode=mode)
    assert len(bits) == int(rec["nbits"])
    assert np.packbits(np.frombuffer(bits, dtype=np.uint8)).tobytes() == rec["packed"].tobytes()
    packed = rec["packed"].tobytes()
    if "dec_error" in rec:
        with pytest.raises(EXC[str(rec["dec_error"][0])]) as ei:
            coder.cabac_decode(packed, coder.ContextModel(n_symbols=n), codes.shape, mode=mode)
        assert "at symbol %d" % int(rec["dec_fault_index"]) in str(ei.value)
        return
    dec = coder.cabac_decode(packed, coder.ContextModel(n_symbols=n), codes.shape, mode=mode)
    assert dec.dtype == np.int32 and np.array_equal(dec, rec["decoded"])


@pytest.mark.parametrize("fixture", ["kat.npz", "coder_small.npz", "coder_full.npz"])
def test_golden_vectors(fixture):
    cases = coder_cases(golden(fixture))
    ran = 0
    for name, rec in cases.items():
        if not _pow2(int(rec["n"])):
            continue
        try:
            _check_case(rec)
        except AssertionError as e:
            raise AssertionError("case %s: %s" % (name, e))
        ran += 1
    assert ran > 100 or fixture != "coder_small.npz"


def test_config1_golden():
    c = golden("config1.npz")
    for mode in ("repaired", "verbatim"):
        rec = {k.split("__", 1)[1]: c[k] for k in c.files if k.startswith("coder_%s__" % mode)}
        rec["mode"] = mode
        _check_case(rec)
    assert hashlib.sha256(c["coder_repaired__packed"].tobytes()).hexdigest()  # fixture present


SETTINGS = [(16, "wide", 24), (16, "enc_like", 24), (256, "enc_like", 48), (256, "wide", 16), (256, "uniform", 16),
            (1024, "enc_like", 12), (64, "hier", 16)]


@pytest.mark.parametrize("n,kind,B", SETTINGS)
def test_batches_vs_oracle(n, kind, B):
    """B independent streams in one launch; every bitstream, every decoded index (including the
    4-bit enc_like streams the reference itself mis-decodes, hazard H1) equals the oracle's."""
    from image_compression_2_b200 import LatentPipeline
    lat = synth_latents(kind, B, 1000 + n)
    pipe = LatentPipeline(n_symbols=n)
    out = pipe.roundtrip_device(lat.cuda())
    idx = out["idx"].cpu().numpy()
    cb = pipe.codebook.cpu().numpy()
    assert np.array_equal(idx, O.quantize_codebook(lat.numpy(), cb))
    streams, nbits, status, fault = out["enc"].to_host()
    dec = out["dec_idx"].cpu().numpy()
    dst = out["dec_status"].cpu().numpy()
    dfi = out["dec_fault"].cpu().numpy()
    deq = out["deq"].cpu().numpy()
    for b in range(B):
        ref = O.encode_stream(idx[b:b + 1], n, "repaired")
        assert status[b] == ref["status"] == 0
        assert nbits[b] == ref["nbits"] and streams[b] == ref["packed"], "stream %d" % b
        rd = O.decode_stream(ref["packed"], n, (1, 16, 512), "repaired")
        assert dst[b] == rd["status"] and dfi[b] == rd["fault_index"], "stream %d" % b
        assert np.array_equal(dec[b], rd["symbols"][0]), "stream %d" % b
        k = rd["fault_index"] if rd["status"] else 16 * 512
        assert np.array_equal(deq[b].ravel()[:k], cb[rd["symbols"][0].ravel()[:k]])


def test_cfg2_full_batch_roundtrip_and_sampled_parity():
    """BASELINE config 2 at full size: 1024 streams of 16x512 at 8 bits. Round trip must be the
    identity on every stream (size-independent property); a seeded sample is compared bit for bit."""
    from image_compression_2_b200 import LatentPipeline
    B, n = 1024, 256
    lat = synth_latents("enc_like", B, 1000 + 2 * 100000)
    pipe = LatentPipeline(n_symbols=n)
    out = pipe.roundtrip_device(lat.cuda())
    idx = out["idx"]
    assert int(out["enc"].status.abs().sum()) == 0 and int(out["dec_status"].abs().sum()) == 0
    assert torch.equal(out["dec_idx"], idx)
    assert torch.equal(out["deq"], pipe.codebook[idx.long()])
    streams, nbits, _, _ = out["enc"].to_host()
    idx_h = idx.cpu().numpy()
    for b in np.random.default_rng(0).choice(B, 24, replace=False):
        ref = O.encode_stream(idx_h[b:b + 1], n, "repaired")
        assert nbits[b] == ref["nbits"] and streams[b] == ref["packed"]
    bps = nbits.astype(np.float64).mean() / 8192
    assert 7.9 < bps < 8.1  # 8.01-8.02 coded bits/symbol on enc_like latents


def test_reference_batched_semantics_shared_model():
    """cabac_encode(data[B,R,C]) is ONE stream whose images share the model (cabac_compression.py:330-337)."""
    from image_compression_2_b200 import coder
    rng = np.random.default_rng(9)
    codes = np.clip(np.round(rng.normal(128, 6, (3, 8, 200))), 0, 255).astype(np.int32)
    bits = coder.cabac_encode(codes, coder.ContextModel(256))
    ref = O.encode_stream(codes, 256, "repaired")
    assert len(bits) == ref["nbits"] and np.array_equal(np.frombuffer(bits, np.uint8), ref["bits"])
    dec = coder.cabac_decode(ref["packed"], coder.ContextModel(256), codes.shape)
    assert np.array_equal(dec, codes)


def test_corrupt_streams_fault_like_the_oracle():
    from image_compression_2_b200 import coder
    rng = np.random.default_rng(12)
    shape = (24, 4, 128)
    streams = []
    for b in range(shape[0]):
        if b % 3 == 0:
            s = bytes(rng.integers(0, 256, 60).astype(np.uint8))
        else:
            c = rng.choice([0, 255], (1, 4, 128)).astype(np.int32)
            s = bytearray(O.encode_stream(c, 256)["packed"])
            if b % 3 == 1:
                s[int(rng.integers(0, len(s)))] ^= 0x10
            s = bytes(s)
        streams.append(s)
    dec, status, fault = coder.cabac_decode_batch(streams, shape)
    for b in range(shape[0]):
        ref = O.decode_stream(streams[b], 256, (1, 4, 128))
        assert status[b] == ref["status"] and fault[b] == ref["fault_index"]
        assert np.array_equal(dec[b], ref["symbols"][0])


@pytest.mark.parametrize("shape,n", [((1, 16, 512), 256), ((1, 4, 128), 256), ((1, 8, 64), 64)])
def test_truncated_streams_read_zeros_like_the_reference(shape, n):
    """The reference reads zero bits past the end of the stream (:260-270).  Valid streams cut at every byte alignment
    -- 0..13 bytes, and 1..9 bytes short of their length -- decode like the oracle (symbols, status, fault index): the
    decoder's bit reader keeps its next word pre-masked and loads words that straddle the end on its rare path.  The
    [16,512] shape takes the kernel specialised for W+ latents, the others the generic instantiation."""
    from image_compression_2_b200 import coder
    rng = np.random.default_rng(n + shape[1])
    codes = np.clip(np.round(rng.normal(n / 2, n * 0.07, shape)), 0, n - 1).astype(np.int32)
    full = bytes(O.encode_stream(codes, n)["packed"])
    cuts = sorted(set(list(range(0, 14)) + [len(full) - k for k in range(0, 10)] + [len(full) // 2, len(full) // 2 + 1,
                                                                                     len(full) // 2 + 2, len(full) // 2 + 3]))
    streams = [full[:c] for c in cuts if 0 <= c <= len(full)]
    dec, status, fault = coder.cabac_decode_batch(streams, (len(streams),) + shape[1:], n_symbols=n)
    for b, s in enumerate(streams):
        ref = O.decode_stream(s, n, shape)
        assert status[b] == ref["status"] and fault[b] == ref["fault_index"], (len(s), status[b], ref["status"], fault[b], ref["fault_index"])
        assert np.array_equal(dec[b], ref["symbols"][0]), len(s)


def _model_of(cm):
    """coder.ContextModel state as the oracle's {(left, up): (vector, count)} (global context: (-2, -2))."""
    return {((-2, -2) if len(k) == 0 else (int(k[0]), int(k[1]))): (np.asarray(v, np.float64), int(cm.context_counts.get(k, 0)))
            for k, v in cm.context_models.items()}


def _same_model(a, b):
    assert set(a) == set(b)
    for k in a:
        assert a[k][1] == b[k][1] and np.array_equal(a[k][0], b[k][0]), k


@pytest.mark.parametrize("n,shape", [(16, (2, 3, 30)), (64, (1, 4, 48)), (256, (1, 16, 128)), (16, (50,)), (1024, (40,)),
                                     (1024, (2, 4, 96))])  # 10 bits with (left,up) contexts: cfg3's alphabet
def test_shared_context_model_across_calls(n, shape):
    """The reference mutates ONE ContextModel in cabac_encode, cabac_decode and across calls (defect D5).  With a
    non-empty model (or track_state=True) the stream is coded from that state and the object ends up holding what
    the reference object would: bits, decoded symbols and every vector/count equal the oracle's (which is pinned
    against the live reference in tests/test_oracle_vs_reference.py)."""
    from image_compression_2_b200 import coder
    rng = np.random.default_rng(5 + n)
    a = np.clip(np.round(rng.normal(n / 2, max(1, n / 16), shape)), 0, n - 1).astype(np.int32)
    b = np.clip(np.round(rng.normal(n / 2, max(1, n / 16), shape)), 0, n - 1).astype(np.int32)
    cm = coder.ContextModel(n, track_state=True)
    model = {}
    for codes in (a, b):
        packed, nbits = coder.cabac_encode_packed(codes, cm)
        ref, model = O.encode_stream_model(codes, n, model)
        assert nbits == ref["nbits"] and packed == ref["packed"]
        _same_model(_model_of(cm), model)
    # a fresh untracked model is left alone and gives the fast kernels' (= fresh-model) stream
    cm0 = coder.ContextModel(n)
    packed0, _ = coder.cabac_encode_packed(a, cm0)
    assert cm0.is_fresh() and packed0 == O.encode_stream(a, n)["packed"]
    # decoding a fresh-model stream with the mutated model (what CABACCompressor.decompress does in the reference)
    ref, model2 = O.decode_stream_model(packed0, n, a.shape, model)
    try:
        dec = coder.cabac_decode(packed0, cm, a.shape)
        assert ref["status"] == 0 and np.array_equal(dec, ref["symbols"])
    except (IndexError, ZeroDivisionError) as e:
        assert ref["status"] != 0, e
    _same_model(_model_of(cm), model2)
    # decoding with the matching model state round-trips
    cm_e, cm_d = coder.ContextModel(n, track_state=True), coder.ContextModel(n, track_state=True)
    for codes in (a, b):
        packed, _ = coder.cabac_encode_packed(codes, cm_e)
        assert np.array_equal(coder.cabac_decode(packed, cm_d, codes.shape), codes)
    _same_model(_model_of(cm_e), _model_of(cm_d))


def test_stateful_model_limits():
    from image_compression_2_b200 import coder
    cm = coder.ContextModel(2048)
    cm.context_models[(0, 0)] = np.ones(2048) / 2048
    with pytest.raises(NotImplementedError):  # alphabets end at 1024 symbols
        coder.cabac_encode(np.zeros((1, 2, 8), np.int32), cm)
    cm = coder.ContextModel(16)
    cm.context_models[(99, 0)] = np.ones(16) / 16
    with pytest.raises(ValueError):
        coder.cabac_encode(np.zeros((1, 2, 8), np.int32), cm)


def test_bad_symbol_raises_index_error():
    from image_compression_2_b200 import coder
    codes = np.zeros((1, 2, 40), np.int32)
    codes[0, 1, 3] = 16
    with pytest.raises(IndexError):
        coder.cabac_encode(codes, coder.ContextModel(16))


# ---- BASELINE.json configurations at full size: size-independent properties + sampled oracle parity ----

def _roundtrip_full(n, kind, B, seed, quantizer="codebook", sample=12):
    from image_compression_2_b200 import LatentPipeline
    lat = synth_latents(kind, B, seed)
    pipe = LatentPipeline(n_symbols=n, quantizer=quantizer)
    out = pipe.roundtrip_device(lat.cuda())
    idx = out["idx"]
    assert int(out["enc"].status.abs().sum()) == 0
    streams, nbits, _, _ = out["enc"].to_host()
    idx_h = idx.cpu().numpy()
    dec_h = out["dec_idx"].cpu().numpy()
    dst = out["dec_status"].cpu().numpy()
    dfi = out["dec_fault"].cpu().numpy()
    for b in np.random.default_rng(seed).choice(B, sample, replace=False):
        ref = O.encode_stream(idx_h[b:b + 1], n, "repaired")
        assert nbits[b] == ref["nbits"] and streams[b] == ref["packed"], "stream %d" % b
        rd = O.decode_stream(ref["packed"], n, (1, 16, 512), "repaired")
        assert dst[b] == rd["status"] and dfi[b] == rd["fault_index"]
        assert np.array_equal(dec_h[b], rd["symbols"][0])
    return out, pipe, nbits


@pytest.mark.parametrize("bits,kind", [(4, "wide"), (8, "enc_like"), (10, "enc_like")])
def test_cfg3_bits_sweep_4096_streams(bits, kind):
    """Config 3: 4096 latents at 4/8/10 bits; the oracle round-trips these, so decode(encode(x)) == x everywhere."""
    n = 1 << bits
    out, pipe, nbits = _roundtrip_full(n, kind, 4096, 1000 + 3 * 100000 + bits)
    assert int(out["dec_status"].abs().sum()) == 0
    assert torch.equal(out["dec_idx"], out["idx"])
    assert torch.equal(out["deq"], pipe.deq_table[out["idx"].long()])
    bps = nbits.astype(np.float64).mean() / 8192
    assert {4: 3.7 < bps < 4.3, 8: 7.9 < bps < 8.1, 10: 9.9 < bps < 10.1}[bits]


def test_cfg3_4bit_enc_like_decoder_parity_up_to_fault():
    """4-bit enc_like latents: the reference decoder itself mis-decodes these (hazard H1); the GPU decoder
    must reproduce the oracle's symbols, fault class and fault index on every sampled stream."""
    out, pipe, _ = _roundtrip_full(16, "enc_like", 1024, 777, sample=48)
    mism = int((out["dec_idx"] != out["idx"]).any(dim=(1, 2)).sum())
    assert mism > 0  # the hazard is real: some streams do not round-trip, exactly as in the reference


def test_cfg5_hier_codes_roundtrip():
    """Config 5: hierarchical multi-scale latents, quantiser B codes (GumbelSoftmaxCompressor.compress output)."""
    out, pipe, _ = _roundtrip_full(256, "hier", 16384 // 8, 1000 + 5 * 100000)  # one GPU's share of 16,384 on 8
    assert int(out["dec_status"].abs().sum()) == 0 and torch.equal(out["dec_idx"], out["idx"])


def test_cfg4_large_batch_roundtrip():
    """Config 4 on one GPU: 8192 streams (one GPU's share of 65,536 on 8), affine quantiser indices."""
    out, pipe, _ = _roundtrip_full(256, "enc_like", 8192, 1000 + 4 * 100000, quantizer="affine", sample=8)
    assert int(out["dec_status"].abs().sum()) == 0 and torch.equal(out["dec_idx"], out["idx"])
    want = O.quantize_affine(synth_latents("enc_like", 8192, 1000 + 4 * 100000)[:8].numpy(), 8)[1]
    assert np.array_equal(out["deq"][:8].cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("build", ["lat", "thr"])
def test_decoder_v2_and_sparse_encoder_special_paths(build, monkeypatch):
    """(build: the decoder kernel's two register budgets -- 8 or 10 resident streams per SM; the host picks by batch
    size, codec.DEFAULT_DECODE_FLAGS forces one.)  The paths the synthetic latents rarely reach: contexts with 7..32 distinct symbols (pool records) and more
    than 32 (decoder: stream redone by the generic kernel; encoder phase A: dense continuation), small alphabets,
    short rows, several images sharing a model -- bitstreams and decoded symbols equal the oracle's."""
    from image_compression_2_b200 import coder
    from image_compression_2_b200 import codec as _codec
    monkeypatch.setattr(_codec, "DEFAULT_DECODE_FLAGS", {"lat": 1, "thr": 2}[build])  # LC_FLAG_DEC_*_BUILD
    rng = np.random.default_rng(31)
    cases = []
    wide = np.zeros((3, 4, 400), np.int32)
    wide[0, :, 0::2] = 7
    wide[0, :, 1::2] = rng.integers(0, 256, (4, 200))          # > 32 distinct symbols in one context
    wide[1, :, 0::2] = 5
    wide[1, :, 1::2] = rng.integers(100, 120, (4, 200))        # 7..20 distinct symbols
    wide[2] = np.clip(np.round(rng.normal(128, 9, (4, 400))), 0, 255)
    cases.append((256, wide))
    for n, shape, sd in ((8, (4, 6, 64), 1.0), (16, (3, 16, 512), 0.6), (128, (3, 8, 200), 9.0), (64, (5, 3, 5), 4.0),
                         (256, (2, 2, 2000), 30.0)):
        cases.append((n, np.clip(np.round(rng.normal(n / 2, sd, shape)), 0, n - 1).astype(np.int32)))
    for n, codes in cases:
        streams, nbits, status, fault = coder.cabac_encode_batch(codes, n_symbols=n)
        for b in range(codes.shape[0]):
            ref = O.encode_stream(codes[b:b + 1], n, "repaired")
            assert status[b] == 0 and nbits[b] == ref["nbits"] and streams[b] == ref["packed"], (n, b)
        dec, dst, dfi = coder.cabac_decode_batch(streams, codes.shape, n_symbols=n)
        for b in range(codes.shape[0]):
            rd = O.decode_stream(streams[b], n, (1,) + codes.shape[1:], "repaired")
            k = int(rd["fault_index"]) if rd["status"] else codes[b].size
            assert dst[b] == rd["status"] and (rd["status"] == 0 or dfi[b] == k), (n, b)
            assert np.array_equal(dec[b].ravel()[:k], rd["symbols"].ravel()[:k]), (n, b)
    # the reference's batched call: images share coder and model
    imgs = np.clip(np.round(rng.normal(32, 3, (3, 4, 64))), 0, 63).astype(np.int32)
    ref = O.encode_stream(imgs, 64, "repaired")
    packed, nb = coder.cabac_encode_packed(imgs, coder.ContextModel(64))
    assert nb == ref["nbits"] and packed == ref["packed"]
    assert np.array_equal(coder.cabac_decode(packed, coder.ContextModel(64), imgs.shape), imgs)
    # a tiny output slot is reported, not overrun (two-kernel phase B)
    from image_compression_2_b200 import codec
    idx = torch.from_numpy(rng.integers(0, 256, (2, 4, 64)).astype(np.int32)).cuda()
    enc = codec.encode_batch(idx.reshape(-1), codec.layout_independent((2, 4, 64)), 256, slot_bytes=64)
    assert enc.status.cpu().tolist() == [5, 5]


def test_decoder_throughput_build_more_streams_than_blocks():
    """1480 + 7 streams of the W+ shape: the host picks the 10-streams-per-SM build and its blocks loop over streams;
    same symbols as the encoder input, and as the 8-per-SM build."""
    from image_compression_2_b200 import codec
    B = 148 * 10 + 7
    g = torch.Generator().manual_seed(5)
    codes = torch.clamp(torch.round(torch.randn(B, 16, 512, generator=g) * 18 + 128), 0, 255).to(torch.int32).cuda()
    layout = codec.layout_independent((B, 16, 512))
    enc = codec.encode_batch(codes.reshape(-1), layout, 256)
    assert not enc.status.any()
    idx, _, status, fault = codec.decode_batch(enc.data, enc.offsets, enc.nbits, layout, 256)
    assert not status.any() and torch.equal(idx.reshape(codes.shape), codes)


def test_symbol_minus_one_is_followed_like_the_reference():
    """The reference does not fault on a decoded symbol -1, it carries on with NumPy's negative indexing
    (cabac_compression.py:288-292,403).  Streams that pass through that state are found with the oracle (which is
    pinned against the live reference on them, tests/test_oracle_vs_reference.py); the kernels must return the same
    array (with its -1 entries), raise the same exception class at the same symbol, and dequantise -1 to
    codebook[-1]."""
    from image_compression_2_b200 import codec, coder
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
        if ref["status"]:
            with pytest.raises(EXC[{1: "ValueError", 2: "IndexError", 3: "ZeroDivisionError"}[ref["status"]]]) as ei:
                coder.cabac_decode(packed, coder.ContextModel(n_symbols=n), shape, mode=mode)
            assert "at symbol %d" % k in str(ei.value)
        else:
            dec = coder.cabac_decode(packed, coder.ContextModel(n_symbols=n), shape, mode=mode)
            assert np.array_equal(dec, ref["symbols"]) and (dec == -1).any()
        # batch form with the fused dequantiser: -1 reads the last table entry
        cb = torch.linspace(-1, 1, n).float()
        data, offsets, nbits = codec.pack_streams_for_device([packed], "cuda")
        idx, deq, st, fi = codec.decode_batch(data, offsets, nbits, codec.layout_independent(shape), n, mode=mode,
                                              codebook=cb.cuda())
        assert int(st.cpu()[0]) == ref["status"]
        assert np.array_equal(idx.cpu().numpy().ravel()[:k], ref["symbols"].ravel()[:k])
        assert np.array_equal(deq.cpu().numpy().ravel()[:k], cb.numpy()[ref["symbols"].ravel()[:k]])
