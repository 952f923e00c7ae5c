"""K1/K2 on the B200 through the C ABI, against the reference's golden vectors and the C oracle. Bit-exact."""
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


def test_quantiser_a_golden():
    from image_compression_2_b200 import codec
    q = golden("quantizers.npz")
    w = torch.from_numpy(q["w"]).cuda()
    for bits in (4, 6, 8, 10):
        idx, wq = codec.quantize_affine(w, bits)
        ref_idx = q["a_idx_%d" % bits]
        fin = np.isfinite(ref_idx) & (np.abs(ref_idx) < 2e9)
        assert np.array_equal(idx.cpu().numpy()[fin], ref_idx[fin].astype(np.int32))
        assert _eq_f32(wq.cpu().numpy(), q["a_wq_%d" % bits])
        good = torch.from_numpy(np.where(fin, ref_idx, 0).astype(np.int32)).cuda()
        assert _eq_f32(codec.dequantize_affine(good, bits).cpu().numpy()[fin], q["a_wq_%d" % bits][fin])


def test_quantiser_b_golden():
    from image_compression_2_b200 import codec
    q = golden("quantizers.npz")
    for n in (16, 64, 256, 1024):
        z = torch.from_numpy(q["b_z_%d" % n]).cuda()
        cb = torch.from_numpy(q["codebook_%d" % n]).cuda()
        for sorted_flag in (True, False):
            idx, deq = codec.quantize_codebook(z, cb, want_deq=True, sorted_ascending=sorted_flag)
            assert np.array_equal(idx.cpu().numpy(), q["b_idx_%d" % n]), (n, sorted_flag)
            assert _eq_f32(deq.cpu().numpy(), q["b_deq_%d" % n])
        assert _eq_f32(codec.dequantize_codebook(idx, cb).cpu().numpy(), q["b_deq_%d" % n])


def test_config1_quantisers():
    from image_compression_2_b200 import codec
    c = golden("config1.npz")
    means = torch.from_numpy(c["means"]).cuda()
    for bits in (4, 8, 10):
        idx, wq = codec.quantize_affine(means, bits)
        assert np.array_equal(idx.cpu().numpy(), c["a_idx_%d" % bits].astype(np.int32))
        assert _eq_f32(wq.cpu().numpy(), c["a_wq_%d" % bits])
    idx, deq = codec.quantize_codebook(means, torch.from_numpy(c["codebook_256"]).cuda(), want_deq=True)
    assert np.array_equal(idx.cpu().numpy(), c["b_idx_256"])
    assert _eq_f32(deq.cpu().numpy(), c["b_deq_256"])


@pytest.mark.parametrize("kind", ["enc_like", "wide", "uniform", "hier"])
def test_quantisers_random_vs_oracle(kind):
    from image_compression_2_b200 import codec
    lat = synth_latents(kind, 64, 4242)
    # ragged length: exercises the non-multiple-of-4 tail
    flat = lat.reshape(-1)[: lat.numel() - 3].contiguous()
    for x in (lat, flat):
        xg = x.cuda()
        for bits in (4, 8, 10):
            n = 1 << bits
            idx, wq = codec.quantize_affine(xg, bits)
            oi, ow = O.quantize_affine(x.numpy(), bits)
            assert np.array_equal(idx.cpu().numpy(), oi) and _eq_f32(wq.cpu().numpy(), ow)
            cb = torch.linspace(-1, 1, n).float()
            bi, bd = codec.quantize_codebook(xg, cb.cuda(), want_deq=True)
            ob = O.quantize_codebook(x.numpy(), cb.numpy())
            assert np.array_equal(bi.cpu().numpy(), ob)
            assert _eq_f32(bd.cpu().numpy(), cb.numpy()[ob])


def test_unsorted_codebook_uses_full_scan():
    from image_compression_2_b200 import codec
    g = torch.Generator().manual_seed(3)
    cb = torch.rand(100, generator=g) * 2 - 1
    z = torch.randn(4096, generator=g) * 0.6
    idx, _ = codec.quantize_codebook(z.cuda(), cb.cuda())
    assert np.array_equal(idx.cpu().numpy(), O.quantize_codebook(z.numpy(), cb.numpy()))


def test_cpu_tensor_is_refused():
    from image_compression_2_b200 import codec
    with pytest.raises(RuntimeError):
        codec.quantize_affine(torch.zeros(16), 8)


@pytest.mark.parametrize("n", [2, 16, 64, 256, 1024])
def test_quantiser_b_decision_boundaries(n):
    """Quantiser B's no-lookup path (near-uniform tables) must hand every value near a decision boundary to the exact
    path: every midpoint of neighbouring entries, +-1..3 ulps and +-1e-3..3e-3 of a step around it, the entries
    themselves, values just outside the table, huge values, specials; int32 / uint16 / uint8 index outputs."""
    from image_compression_2_b200 import codec
    cb = torch.linspace(-1, 1, n).float().numpy()
    mid = ((cb[:-1].astype(np.float64) + cb[1:].astype(np.float64)) / 2).astype(np.float32)
    step = np.float32(2.0 / (n - 1))
    pts = [cb, mid]
    for k in (1, 2, 3):
        up, dn = mid.copy(), mid.copy()
        for _ in range(k):
            up = np.nextafter(up, np.float32(np.inf)); dn = np.nextafter(dn, np.float32(-np.inf))
        pts += [up, dn]
    for f in (1e-3, 1.9e-3, 2.1e-3, 3e-3, 0.4, 0.49, 0.499):
        pts += [mid + np.float32(f) * step, mid - np.float32(f) * step, cb + np.float32(f) * step, cb - np.float32(f) * step]
    pts.append(np.array([-1.0000001, 1.0000001, -1.5, 1.5, -7.9, 7.9, -8.1, 8.1, 1e9, -1e9, 3e38, -3e38, np.inf, -np.inf,
                         np.nan, 0.0, -0.0, 1e-30], np.float32))
    rng = np.random.default_rng(n)
    pts.append(rng.uniform(-1.2, 1.2, 20000).astype(np.float32))
    z = np.concatenate(pts).astype(np.float32)
    z = np.concatenate([z, z[: (-len(z)) % 4 + 1]])  # ragged tail
    ref = O.quantize_codebook(z, cb)
    zg, cbg = torch.from_numpy(z).cuda(), torch.from_numpy(cb).cuda()
    for dt in (torch.int32, torch.int16, torch.uint8):
        if dt == torch.uint8 and n > 256:
            continue
        idx, deq = codec.quantize_codebook(zg, cbg, want_deq=True, idx_dtype=dt)
        got = idx.cpu().numpy().astype(np.int64) & (0xFFFF if dt == torch.int16 else 0xFFFFFFFF if dt == torch.int32 else 0xFF)
        assert np.array_equal(got, ref.astype(np.int64)), (n, dt, np.flatnonzero(got != ref)[:8], z[got != ref][:8])
        assert _eq_f32(deq.cpu().numpy(), cb[ref])


def test_dequantiser_a_table_and_out_of_table_indices():
    """Dequantiser A tabulates the 2^bits values per block; int32 indices outside the table take the arithmetic."""
    from image_compression_2_b200 import codec
    rng = np.random.default_rng(5)
    for bits in (1, 4, 8, 10, 12, 13, 16):
        hi = (1 << bits) - 1
        idx = np.concatenate([np.arange(0, min(hi, 5000) + 1), rng.integers(0, hi + 1, 10001),
                              np.array([-1, -5, hi + 1, hi + 100, 1 << 20, -(1 << 20)])]).astype(np.int32)
        got = codec.dequantize_affine(torch.from_numpy(idx).cuda(), bits).cpu().numpy()
        assert _eq_f32(got, O.dequantize_affine(idx, bits)), bits


def test_quantiser_b_sorted_non_uniform_tables():
    """A trained codebook is sorted but not a linspace: the two-lookup path (well-separated entries) and the
    three-lookup / exact-search path (entries closer than 1e-5, duplicates) against the oracle."""
    from image_compression_2_b200 import codec
    g = torch.Generator().manual_seed(11)
    z = torch.cat([torch.randn(50001, generator=g) * 0.5, torch.tensor([-3.0, 3.0, 0.0, float("nan"), float("inf"), -float("inf")])])
    base = torch.sort(torch.rand(256, generator=g) * 2 - 1).values
    close = base.clone()
    close[100] = close[99] + 2e-6      # closer than the separation the two-lookup path needs
    close[200] = close[199]            # a duplicate entry: argmin must return the first
    close = torch.sort(close).values
    warped = torch.sort(torch.tanh(torch.linspace(-2, 2, 64))).values  # smooth but far from uniform
    for cb in (base, close, warped):
        ref = O.quantize_codebook(z.numpy(), cb.numpy())
        for dt in (torch.int32, torch.uint8):
            idx, deq = codec.quantize_codebook(z.cuda(), cb.cuda(), want_deq=True, idx_dtype=dt)
            assert np.array_equal(idx.cpu().numpy().astype(np.int64), ref.astype(np.int64))
            assert _eq_f32(deq.cpu().numpy(), cb.numpy()[ref])


@pytest.mark.parametrize("amp", [1e-3, 0.02, 0.03])
def test_quantiser_b_nearly_uniform_tables(amp):
    """A linspace whose entries are displaced by up to `amp` steps: the no-lookup path widens its exact band to the
    table's measured deviation (amp = 0.03 is past its limit and takes the two-lookup path); dense sampling around every
    decision boundary."""
    from image_compression_2_b200 import codec
    n = 256
    rng = np.random.default_rng(int(amp * 1e6))
    step = 2.0 / (n - 1)
    cb = (np.linspace(-1, 1, n) + rng.uniform(-amp, amp, n) * step).astype(np.float32)
    cb[0], cb[-1] = -1.0, 1.0
    assert (np.diff(cb) > 0).all()
    mid = ((cb[:-1].astype(np.float64) + cb[1:].astype(np.float64)) / 2).astype(np.float32)
    pts = [rng.uniform(-1.1, 1.1, 40000).astype(np.float32), cb, mid]
    for f in np.linspace(-2.5 * amp - 1e-3, 2.5 * amp + 1e-3, 41):
        pts.append(mid + np.float32(f * step))
    z = np.concatenate(pts).astype(np.float32)
    ref = O.quantize_codebook(z, cb)
    idx, _ = codec.quantize_codebook(torch.from_numpy(z).cuda(), torch.from_numpy(cb).cuda(), idx_dtype=torch.uint8)
    got = idx.cpu().numpy().astype(np.int64)
    assert np.array_equal(got, ref.astype(np.int64)), (amp, np.flatnonzero(got != ref)[:8], z[got != ref][:8])
