"""ABI version 2 on the B200: typed (uint8 / uint16) index arrays, the `flags` argument, per-device state (two devices
from one process, two host threads on one device), ContextModel's methods, the overflow retry of the drop-in class."""
import threading

import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.helpers import golden, synth_latents

pytestmark = pytest.mark.gpu


def _eq_f32(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.array_equal(a.view(np.uint32)[~np.isnan(a)], b.view(np.uint32)[~np.isnan(b)]) and \
        np.array_equal(np.isnan(a), np.isnan(b))


@pytest.mark.parametrize("dtype", [torch.uint8, torch.int16])
def test_narrow_index_quantisers_match_the_int32_forms(dtype):
    """uint8 / uint16 indices = the int32 indices clamped to the alphabet; fp32 outputs are bit-identical."""
    from image_compression_2_b200 import codec
    q = golden("quantizers.npz")
    w = torch.from_numpy(q["w"]).cuda()
    for bits in ((4, 6, 8) if dtype == torch.uint8 else (4, 8, 10)):
        i32, wq32 = codec.quantize_affine(w, bits)
        i8, wq8 = codec.quantize_affine(w, bits, idx_dtype=dtype)
        ref = q["a_idx_%d" % bits]
        want = np.where(np.isnan(ref), 0, np.clip(ref, 0, (1 << bits) - 1)).astype(np.int64)
        assert np.array_equal(i8.cpu().numpy().astype(np.int64), want), bits
        assert _eq_f32(wq8.cpu().numpy(), wq32.cpu().numpy())
        # dequantiser A on the narrow type = on int32
        assert _eq_f32(codec.dequantize_affine(i8, bits).cpu().numpy(),
                       codec.dequantize_affine(i8.to(torch.int32), bits).cpu().numpy())
    for n in ((16, 64, 256) if dtype == torch.uint8 else (16, 256, 1024)):
        z = torch.from_numpy(q["b_z_%d" % n]).cuda()
        cb = torch.from_numpy(q["codebook_%d" % n]).cuda()
        idx, deq = codec.quantize_codebook(z, cb, want_deq=True, idx_dtype=dtype)
        assert np.array_equal(idx.cpu().numpy().astype(np.int64), q["b_idx_%d" % n].astype(np.int64)), n
        assert _eq_f32(deq.cpu().numpy(), q["b_deq_%d" % n])
        assert _eq_f32(codec.dequantize_codebook(idx, cb).cpu().numpy(), q["b_deq_%d" % n])
    # odd lengths exercise the scalar tail of the vectorised kernels
    z = torch.from_numpy(q["b_z_256"]).cuda().reshape(-1)[:1003].contiguous()
    cb = torch.from_numpy(q["codebook_256"]).cuda()
    a, _ = codec.quantize_codebook(z, cb, idx_dtype=dtype)
    b, _ = codec.quantize_codebook(z, cb)
    assert torch.equal(a.to(torch.int32), b)
    with pytest.raises(ValueError):  # 1024 symbols do not fit a byte
        codec.quantize_codebook(z, torch.linspace(-1, 1, 1024).cuda(), idx_dtype=torch.uint8)


@pytest.mark.parametrize("bits,kind", [(4, "wide"), (8, "enc_like"), (10, "enc_like")])
def test_coder_on_narrow_indices_matches_the_oracle(bits, kind):
    """uint8 / uint16 codes in, uint8 / uint16 symbols out: same bitstreams as int32 codes and as the oracle."""
    from image_compression_2_b200 import codec
    n = 1 << bits
    B = 6
    lat = synth_latents(kind, B, 77 + bits)
    cb = torch.linspace(-1, 1, n).float()
    idx32 = O.quantize_codebook(lat.numpy(), cb.numpy())
    layout = codec.layout_independent(idx32.shape)
    dt = codec.idx_dtype_for(n)
    narrow = torch.from_numpy(idx32).to(dt).cuda()
    enc_n = codec.encode_batch(narrow.reshape(-1), layout, n)
    enc_w = codec.encode_batch(torch.from_numpy(idx32).cuda().reshape(-1), layout, n)
    sn, nbn, stn, _ = enc_n.to_host()
    sw, nbw, stw, _ = enc_w.to_host()
    assert not stn.any() and not stw.any()
    for b in range(B):
        ref = O.encode_stream(idx32[b:b + 1], n, "repaired")
        assert nbn[b] == nbw[b] == ref["nbits"] and sn[b] == sw[b] == ref["packed"], b
    for want_idx in (True, False):
        idx, deq, st, _ = codec.decode_batch(enc_n.data, enc_n.offsets, enc_n.nbits, layout, n, codebook=cb.cuda(),
                                             idx_dtype=dt, want_idx=want_idx)
        assert not st.cpu().numpy().any()
        assert (idx is None) == (not want_idx)
        if want_idx:
            assert idx.dtype == dt and np.array_equal(idx.cpu().numpy().astype(np.int64).reshape(idx32.shape), idx32)
        assert _eq_f32(deq.cpu().numpy().reshape(idx32.shape), cb.numpy()[idx32])


def test_flags_select_kernels_without_changing_results():
    from image_compression_2_b200 import _native, codec
    n, B = 256, 5
    lat = synth_latents("enc_like", B, 4242)
    idx32 = O.quantize_codebook(lat.numpy(), torch.linspace(-1, 1, n).numpy())
    layout = codec.layout_independent(idx32.shape)
    dev_idx = torch.from_numpy(idx32).cuda().reshape(-1)
    refs = [O.encode_stream(idx32[b:b + 1], n, "repaired") for b in range(B)]
    for eflags in (0, _native.FLAG_ENC_SERIAL, _native.FLAG_ENC_SORT_V1):
        enc = codec.encode_batch(dev_idx, layout, n, flags=eflags)
        streams, nbits, status, _ = enc.to_host()
        assert not status.any()
        for b in range(B):
            assert nbits[b] == refs[b]["nbits"] and streams[b] == refs[b]["packed"], (eflags, b)
    for dflags in (0, _native.FLAG_DEC_LATENCY_BUILD, _native.FLAG_DEC_THROUGHPUT_BUILD,
                   _native.FLAG_DEC_GENERIC_SHAPE, _native.FLAG_DEC_GENERIC_SHAPE | _native.FLAG_DEC_THROUGHPUT_BUILD,
                   _native.FLAG_DEC_REGISTER_MODEL, _native.FLAG_DEC_SERIAL):
        idx, _, st, _ = codec.decode_batch(enc.data, enc.offsets, enc.nbits, layout, n, flags=dflags)
        assert not st.cpu().numpy().any(), dflags
        assert np.array_equal(idx.cpu().numpy().reshape(idx32.shape), idx32), dflags


def test_launch_counter_counts_this_threads_launches():
    from image_compression_2_b200 import _native, codec
    lib = _native.load()
    w = torch.zeros(4096, device="cuda")
    lib.lc_debug_launch_count(1)
    codec.quantize_affine(w, 8)
    codec.quantize_affine(w, 8)
    assert lib.lc_debug_launch_count(1) == 2
    assert lib.lc_debug_launch_count(0) == 0


def test_two_host_threads_share_one_device():
    """No mutable library state: two threads, each on its own CUDA stream and workspace, code different batches
    concurrently and both match the oracle."""
    from image_compression_2_b200 import codec
    n = 256
    results, errors = {}, []

    def work(tid):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for rep in range(3):
                    lat = synth_latents("enc_like", 24, 900 + 10 * tid + rep)
                    idx32 = O.quantize_codebook(lat.numpy(), torch.linspace(-1, 1, n).numpy())
                    layout = codec.layout_independent(idx32.shape)
                    enc = codec.encode_batch(torch.from_numpy(idx32).cuda().reshape(-1), layout, n)
                    idx, _, dst, _ = codec.decode_batch(enc.data, enc.offsets, enc.nbits, layout, n)
                    st.synchronize()
                    streams, nbits, status, _ = enc.to_host()
                    ok = not status.any() and not dst.cpu().numpy().any() and \
                        np.array_equal(idx.cpu().numpy().reshape(idx32.shape), idx32)
                    for b in (0, 11, 23):
                        ref = O.encode_stream(idx32[b:b + 1], n, "repaired")
                        ok = ok and nbits[b] == ref["nbits"] and streams[b] == ref["packed"]
                    results[(tid, rep)] = ok
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert len(results) == 6 and all(results.values()), results


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_from_one_process():
    """Per-device caches: the first call on a second device must set that device's kernel attributes and size its
    grids from that device -- tensors on cuda:1 while cuda:0 is the current device."""
    from image_compression_2_b200 import LatentPipeline, codec
    n = 256
    lat = synth_latents("enc_like", 16, 31337)
    idx32 = O.quantize_codebook(lat.numpy(), torch.linspace(-1, 1, n).numpy())
    torch.cuda.set_device(0)
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        pipe = LatentPipeline(n_symbols=n, device=dev)
        out = pipe.roundtrip_device(lat.to(dev))
        torch.cuda.synchronize(dev)
        assert out["idx"].device == torch.device(dev)
        assert np.array_equal(out["idx"].cpu().numpy(), idx32)
        assert not out["dec_status"].cpu().numpy().any() and torch.equal(out["dec_idx"], out["idx"])
        streams, nbits, status, _ = out["enc"].to_host()
        ref = O.encode_stream(idx32[3:4], n, "repaired")
        assert nbits[3] == ref["nbits"] and streams[3] == ref["packed"]
    enc = out["enc"]
    with pytest.raises(RuntimeError):  # operands on two devices are refused, not silently launched on one of them
        codec.decode_batch(enc.data, enc.offsets.to("cuda:1"), enc.nbits, enc.layout, n)


def _np_update(p, s, rate=0.05):
    """ContextModel.update_model as the reference writes it (cabac_compression.py:119-144), NumPy on the host."""
    q = p.copy()
    q[s] += rate * (1.0 - q[s])
    total_others = q.sum() - q[s]
    f = (1.0 - q[s]) / total_others if total_others > 0 else 0
    for i in range(len(q)):
        if i != s:
            q[i] *= f
    return q


@pytest.mark.parametrize("n", [4, 16, 256, 1024])
def test_context_model_methods(n):
    from image_compression_2_b200 import ContextModel
    cm = ContextModel(n_symbols=n)
    data = np.arange(24, dtype=np.int32).reshape(2, 3, 4) % n
    assert cm.get_context(data, 0, data.shape) == (-1, -1)
    assert cm.get_context(data, 5, data.shape) == (int(data[0, 1, 0]), int(data[0, 0, 1]))
    assert cm.get_context(data, 12, data.shape) == (-1, -1)  # first element of the second image
    assert cm.get_context(data.reshape(-1), 7, (24,)) == ()
    p0 = cm.get_probability((1, 2))
    assert np.array_equal(p0, np.ones(n) / n) and (1, 2) in cm.context_models
    rng = np.random.default_rng(n)
    want = p0.copy()
    for s in [int(x) for x in rng.integers(0, n, 5)] + [-1, n - 1, -n]:
        cm.update_model((1, 2), s)
        want = _np_update(want, s)
        assert np.array_equal(cm.context_models[(1, 2)], want), s
    assert cm.context_counts[(1, 2)] == 8
    assert cm.get_probability((1, 2), 3) == want[3]
    with pytest.raises(IndexError):
        cm.update_model((1, 2), n)


def test_compress_retries_streams_that_overflow_the_default_slot():
    """A stream that needs far more than log2(n)+2 bits per symbol (every symbol is the least likely one of its
    context) still compresses through the drop-in class, as it does in the reference."""
    from image_compression_2_b200 import codec, coder
    n = 16
    # alternate between two symbols per context so each context always predicts the other one; add a ramp
    codes = np.zeros((1, 4, 256), np.int32)
    codes[0, :, :] = (np.arange(256)[None, :] * 7 + np.arange(4)[:, None] * 3) % n
    layout = codec.layout_reference(codes.shape)
    tiny = 64  # bytes: certainly too small
    enc = codec.encode_batch(torch.from_numpy(codes).cuda().reshape(-1), layout, n, slot_bytes=tiny)
    assert int(enc.status.cpu()[0]) == 5
    _, (streams, nbits, status, _) = codec.encode_batch_checked(torch.from_numpy(codes).cuda().reshape(-1), layout, n,
                                                                slot_bytes=tiny)
    ref = O.encode_stream(codes, n, "repaired")
    assert status[0] == 0 and nbits[0] == ref["nbits"] and streams[0] == ref["packed"]
    packed, nb = coder.cabac_encode_packed(codes, coder.ContextModel(n))
    assert nb == ref["nbits"] and packed == ref["packed"]


@pytest.mark.parametrize("n,sd", [(16, 1.14), (16, 3.0), (8, 1.0), (4, 0.8), (2, 0.5)])
def test_small_alphabet_decoder_and_its_alternatives_agree_with_the_oracle(n, sd):
    """Alphabets up to 16 symbols take the dense shared-memory decoder (lc_decoder_small.cuh); LC_FLAG_DEC_NO_SMALL
    sends them down the other kernels.  Narrow 4-bit data does not round-trip in the reference (hazards H1-H3): the
    test is decoder-vs-decoder -- same symbols up to the fault, same fault class, same fault index."""
    from image_compression_2_b200 import _native, codec
    rng = np.random.default_rng(n * 100 + int(sd * 10))
    shape = (6, 16, 512)
    codes = np.clip(np.round(rng.normal(n / 2, sd, shape)), 0, n - 1).astype(np.int32)
    streams = [O.encode_stream(codes[b:b + 1], n)["packed"] for b in range(shape[0])]
    bad = bytearray(streams[-1]); bad[len(bad) // 3] ^= 0x11; streams[-1] = bytes(bad)   # a corrupted stream
    layout = codec.layout_independent(shape)
    data, offsets, nbits = codec.pack_streams_for_device(streams, "cuda")
    cb = torch.linspace(-1, 1, n).float().cuda()
    refs = [O.decode_stream(streams[b], n, (1,) + shape[1:]) for b in range(shape[0])]
    for flags in (0, _native.FLAG_DEC_NO_SMALL, _native.FLAG_DEC_SERIAL):
        idx, deq, st, fi = codec.decode_batch(data, offsets, nbits, layout, n, codebook=cb, flags=flags,
                                              idx_dtype=torch.uint8)
        idx, deq, st, fi = idx.cpu().numpy(), deq.cpu().numpy(), st.cpu().numpy(), fi.cpu().numpy()
        for b, ref in enumerate(refs):
            k = int(ref["fault_index"]) if ref["status"] else codes[b].size
            assert st[b] == ref["status"] and (ref["status"] == 0 or fi[b] == k), (flags, b, st[b], ref["status"])
            assert np.array_equal(idx[b][:k], ref["symbols"].ravel()[:k]), (flags, b)
            assert np.array_equal(deq[b][:k], cb.cpu().numpy()[ref["symbols"].ravel()[:k]]), (flags, b)
