"""The drop-in compressor classes (reference Python surface) on the GPU, with stub encoder/generator."""
import pickle
import struct

import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.stubs import StubEncoder, StubGenerator

pytestmark = pytest.mark.gpu


def _x(B):
    return torch.zeros(B, 3, 256, 256, device="cuda")


def test_stylegan3_compressor_npz(tmp_path):
    from image_compression_2_b200 import StyleGAN3Compressor
    enc, gen = StubEncoder().cuda(), StubGenerator().cuda()
    comp = StyleGAN3Compressor(enc, gen)
    x = _x(2)
    for bits in (4, 8, 10):
        wq = comp.compress(x, quantization_bits=bits)
        assert wq.shape == (2, 16, 512) and wq.dtype == torch.float32 and wq.is_cuda
        _, ref = O.quantize_affine(enc.means_for(2).numpy(), bits)
        assert np.array_equal(wq.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    fn = str(tmp_path / "lat")
    ret = comp.save_compressed(x, fn, quantization_bits=8)
    assert ret == (1572864, 16384.0, 96.0)  # B=2 @256x256, 8 bits
    data = np.load(fn + ".npz")
    assert sorted(data.files) == ["bits", "comp_size", "compression_ratio", "orig_size", "resolution", "w"]
    assert data["w"].dtype == np.float32 and data["w"].shape == (2, 16, 512)
    assert data["resolution"].dtype == np.int64 and data["resolution"].tolist() == [256, 256]
    assert data["bits"].dtype == np.int64 and data["orig_size"].dtype == np.int64
    assert data["comp_size"].dtype == np.float64 and data["compression_ratio"].dtype == np.float64
    img, ratio = comp.load_compressed(fn + ".npz")
    assert float(ratio) == 96.0 and img.shape[0] == 2
    assert torch.equal(img, gen.synthesis(comp.compress(x, 8)))


def test_gumbel_compressor_codes_and_npz(tmp_path):
    from image_compression_2_b200 import GumbelSoftmaxCompressor
    enc, gen = StubEncoder().cuda(), StubGenerator().cuda()
    comp = GumbelSoftmaxCompressor(enc, gen, n_embeddings=256).cuda()
    x = _x(2)
    codes = comp.compress(x)
    assert codes.dtype == torch.int64 and codes.device.type == "cpu" and codes.shape == (2, 16, 512)
    cb = torch.linspace(-1, 1, 256).float().numpy()
    assert np.array_equal(codes.numpy(), O.quantize_codebook(enc.means_for(2).numpy(), cb))
    fn = str(tmp_path / "codes")
    ret = comp.save_compressed(x, fn)
    assert ret == (1572864, 16384.0, 96.0)
    data = np.load(fn + ".npz")
    assert sorted(data.files) == ["codes", "comp_size", "compression_ratio", "n_embeddings", "orig_size", "resolution"]
    assert data["codes"].dtype == np.int64 and data["n_embeddings"].dtype == np.int64
    img, ratio = comp.load_compressed(fn + ".npz")
    w = torch.from_numpy(cb[codes.numpy()]).cuda()
    assert torch.equal(img, gen.synthesis(w)) and float(ratio) == 96.0
    deq, perplexity, flat = comp.discretization(enc.means_for(2).cuda())
    assert torch.equal(flat.cpu(), codes.reshape(-1)) and float(perplexity) > 1


def test_cabac_compressor_packed_container_roundtrip(tmp_path):
    from image_compression_2_b200 import CABACCompressor
    enc, gen = StubEncoder().cuda(), StubGenerator().cuda()
    comp = CABACCompressor(enc, gen, n_embeddings=256)
    comp.discretization.cuda()
    x = _x(1)
    encoded, meta = comp.compress(x)
    cb = torch.linspace(-1, 1, 256).float().numpy()
    codes = O.quantize_codebook(enc.means_for(1).numpy(), cb)
    ref = O.encode_stream(codes, 256, "repaired")
    assert encoded == ref["packed"]
    assert meta["shape"] == (1, 16, 512) and meta["n_embeddings"] == 256 and meta["use_cabac"] is True
    assert meta["orig_size"] == 8192.0 and meta["comp_size"] == len(ref["packed"])
    fn = str(tmp_path / "img.cabac")
    stats = comp.save_compressed(x, fn)
    assert stats == (meta["orig_size"], meta["comp_size"], meta["compression_ratio"])
    img, ratio = comp.load_compressed(fn)
    assert torch.equal(img, gen.synthesis(torch.from_numpy(cb[codes]).cuda())) and ratio == meta["compression_ratio"]
    # use_cabac=False: raw int32 codes, as in the reference
    raw, meta2 = comp.compress(x, use_cabac=False)
    assert raw == codes.astype(np.int32).tobytes() and meta2["comp_size"] == codes.size * 4
    assert torch.equal(comp.decompress(raw, meta2), img)


def test_cabac_compressor_reference_container_bytes(tmp_path):
    """container='reference' writes exactly what cabac_compression.py:555-561 writes (defects D1, D4)."""
    from image_compression_2_b200 import CABACCompressor
    enc, gen = StubEncoder(seed=3).cuda(), StubGenerator().cuda()
    comp = CABACCompressor(enc, gen, n_embeddings=256, container="reference")
    comp.discretization.cuda()
    x = _x(1)
    fn = str(tmp_path / "ref.cabac")
    orig, comp_size, ratio = comp.save_compressed(x, fn)
    cb = torch.linspace(-1, 1, 256).float().numpy()
    ref = O.encode_stream(O.quantize_codebook(enc.means_for(1).numpy(), cb), 256, "repaired")
    assert comp_size == ref["nbits"]  # counts bits as bytes, like the reference
    blob = open(fn, "rb").read()
    assert struct.unpack("I", blob[:4])[0] == 6
    meta = {"shape": (1, 16, 512), "n_embeddings": 256, "use_cabac": True, "orig_size": orig, "comp_size": comp_size,
            "compression_ratio": ratio}
    pk = pickle.dumps(meta)
    assert blob[4:4 + len(pk)] == pk
    assert blob[4 + len(pk):] == ref["bits"].tobytes()


def test_pipeline_host_roundtrip():
    from image_compression_2_b200 import LatentPipeline
    from tests.helpers import synth_latents
    B = 32
    lat = synth_latents("enc_like", B, 77).pin_memory()
    for quantizer, chunks in (("codebook", None), ("affine", None), ("codebook", 3)):  # 3: uneven chunks on 3 streams
        pipe = LatentPipeline(n_symbols=256, quantizer=quantizer)
        for attempt in range(3):
            # 0: no size hint yet (whole slots travel); 1: hint from the first call; 2: a hint that is too small --
            # the chunks are redone with the exact size.  Same result every time.
            if attempt == 2:
                pipe._bytes_hint = 100
            res = pipe.roundtrip_host(lat, chunks=chunks)
            assert res["chunks"] == (chunks or 1)
            assert not res["enc_status"].numpy().any() and not res["dec_status"].numpy().any()
            if attempt == 1:
                assert res["h2d_bytes"] < first_h2d  # the hint shrinks the compressed-bytes round trip
            first_h2d = res["h2d_bytes"] if attempt == 0 else first_h2d
        if quantizer == "codebook":
            cb = pipe.codebook.cpu().numpy()
            idx = O.quantize_codebook(lat.numpy(), cb)
            want = cb[idx]
        else:
            idx, want = O.quantize_affine(lat.numpy(), 8)
        assert np.array_equal(res["deq"].numpy().view(np.uint32), want.view(np.uint32))
        offs, nbits = res["offsets"].numpy(), res["nbits"].numpy()
        blob = res["bytes"].numpy()
        for b in range(0, B, 5):
            ref = O.encode_stream(idx[b:b + 1], 256, "repaired")
            assert nbits[b] == ref["nbits"]
            assert blob[offs[b]:offs[b] + len(ref["packed"])].tobytes() == ref["packed"]
        assert res["h2d_bytes"] > lat.numel() * 4 and res["d2h_bytes"] > lat.numel() * 4


def test_pipeline_host_stream_of_batches():
    """roundtrip_host_stream: different batches in flight at once (two and three slots; a size hint that is too small)
    give, batch by batch, what the oracle says."""
    from image_compression_2_b200 import LatentPipeline
    from tests.helpers import synth_latents
    kinds = ["enc_like", "wide", "enc_like", "uniform", "hier", "enc_like", "wide"]
    batches = [synth_latents(k, 24 if i % 2 else 32, 900 + i).pin_memory() for i, k in enumerate(kinds)]
    for depth, hint in ((2, 0), (3, 0), (2, 100)):
        pipe = LatentPipeline(n_symbols=256)
        pipe._bytes_hint = hint  # 100: too small for any stream -- the batches in flight are redone one at a time
        cb = pipe.codebook.cpu().numpy()
        seen = 0
        for lat, res in zip(batches, pipe.roundtrip_host_stream(batches, depth=depth)):
            idx = O.quantize_codebook(lat.numpy(), cb)
            assert not res["enc_status"].numpy().any() and not res["dec_status"].numpy().any()
            assert np.array_equal(res["deq"].numpy().view(np.uint32), cb[idx].view(np.uint32)), (depth, seen)
            offs, nbits, blob = res["offsets"].numpy(), res["nbits"].numpy(), res["bytes"].numpy()
            for b in (0, lat.shape[0] - 1):
                ref = O.encode_stream(idx[b:b + 1], 256, "repaired")
                assert nbits[b] == ref["nbits"]
                assert blob[offs[b]:offs[b] + len(ref["packed"])].tobytes() == ref["packed"]
            seen += 1
        assert seen == len(batches)


def test_bitrate_stats_of_a_coded_batch():
    from image_compression_2_b200 import LatentPipeline, stats
    from tests.helpers import synth_latents
    lat = synth_latents("enc_like", 32, 4242)
    pipe = LatentPipeline(n_symbols=256)
    enc = pipe.encode(pipe.quantize(lat.cuda()))
    st = stats.encoded_batch_stats(enc)
    assert st["streams"] == 32 and 7.9 < st["coded_bits_per_symbol"] < 8.1
    assert st["packed_bytes"]["total"] == int(((enc.nbits.cpu().numpy() + 7) // 8).sum())


def test_cabac_compressor_shared_model_like_the_reference():
    """shared_model=True: one ContextModel across compress calls (cabac_compression.py:438,478): the second image is
    coded from the first image's final model -- the oracle's stateful encoder gives the same bytes."""
    from image_compression_2_b200 import CABACCompressor
    enc, gen = StubEncoder().cuda(), StubGenerator().cuda()
    comp = CABACCompressor(enc, gen, n_embeddings=64, shared_model=True)
    fresh = CABACCompressor(enc, gen, n_embeddings=64)
    x1, x2 = torch.randn(1, 3, 64, 64).cuda(), torch.randn(1, 3, 64, 64).cuda()
    model = {}
    for x in (x1, x2):
        encoded, meta = comp.compress(x)
        codes = fresh._codes_device(x).cpu().numpy()
        ref, model = O.encode_stream_model(codes, 64, model)
        assert encoded == ref["packed"] and meta["comp_size"] == len(ref["packed"])
    assert len(comp.context_model.context_models) == len(model)
    e1, _ = fresh.compress(x2)
    assert e1 == O.encode_stream(fresh._codes_device(x2).cpu().numpy(), 64)["packed"] and e1 != encoded


@pytest.mark.parametrize("container", ["packed", "reference"])
def test_compare_compression_methods_report(container):
    """The reference's per-method comparison (cabac_compression.py:800-881) over GPU outputs: same keys, the sizes the
    reference would report (the 'reference' container counts bits as bytes, defect D1), ratios consistent."""
    from image_compression_2_b200 import CABACCompressor, stats
    enc, gen = StubEncoder().cuda(), StubGenerator().cuda()
    comp = CABACCompressor(enc, gen, n_embeddings=256, container=container)
    x = (torch.rand(1, 3, 64, 64, device="cuda") * 2 - 1)
    rep = stats.compare_compression_methods(comp, x)
    assert {"original", "png", "jpg", "hvae", "cabac", "hvae_ratio", "cabac_ratio", "cabac_vs_hvae"} <= set(rep)
    cb = torch.linspace(-1, 1, 256).float().numpy()
    ref = O.encode_stream(O.quantize_codebook(enc.means_for(1).numpy(), cb), 256, "repaired")
    assert rep["hvae"] == 16 * 512 * 4                                    # the int32 codes of use_cabac=False (:484)
    assert rep["cabac"] == (ref["nbits"] if container == "reference" else len(ref["packed"]))
    assert rep["cabac_vs_hvae"] == rep["hvae"] / rep["cabac"]
    assert abs(rep["coded_bits_per_symbol"] - ref["nbits"] / 8192) < 8 / 8192
    assert rep["original"] == 3 * 64 * 64 and len(rep["table"]) == 5
    if rep["png"] is not None:
        assert rep["png"] > 0 and rep["jpg"] > 0
    # the batch form: one table for a whole coded batch
    from image_compression_2_b200 import LatentPipeline
    from tests.helpers import synth_latents
    pipe = LatentPipeline(n_symbols=256)
    enc_b = pipe.encode(pipe.quantize(synth_latents("enc_like", 16, 99).cuda()))
    tab = stats.method_table(enc_b)
    assert tab["streams"] == 16 and [r["method"][:5] for r in tab["rows"]] == ["int32", "fixed", "arith"]
    assert abs(tab["rows"][1]["ratio"] - 4.0) < 1e-9 and 3.9 < tab["rows"][2]["ratio"] < 4.1


def test_gumbel_discretization_loads_a_reference_state_dict():
    """Same parameter/buffer names as the reference layer (gumbel_softmax_compression.py:49-63), so the
    `discretization_state_dict` of a reference checkpoint loads (cabac_compression.py:660-666)."""
    from image_compression_2_b200 import GumbelSoftmaxDiscretization
    ref_state = {"codebook": torch.linspace(-1, 1, 64).float(), "log_temperature": torch.ones(1) * np.log(0.7),
                 "usage": torch.arange(64).float()}
    for learnable in (True, False):
        d = GumbelSoftmaxDiscretization(latent_dim=512, n_embeddings=64, learnable_temp=learnable).cuda()
        assert set(d.state_dict()) == set(ref_state)
        d.load_state_dict(ref_state)
        assert abs(float(d.temperature) - 0.7) < 1e-6
        assert torch.allclose(d.get_code_usage().cpu(), ref_state["usage"] / ref_state["usage"].sum())
        d.update_temp(anneal_rate=0.1, min_temp=0.5)
        assert float(d.log_temperature) < np.log(0.7)
    d.train()
    z = (torch.randn(2, 16, 512, device="cuda") * 0.3)
    before = d.usage.clone()
    deq, perplexity, idx = d(z)
    assert d.usage.sum() == before.sum() + z.numel() and deq.shape == z.shape and idx.shape == (z.numel(),)
    assert np.array_equal(idx.cpu().numpy(), O.quantize_codebook(z.cpu().numpy(), d.codebook.cpu().numpy()).reshape(-1))
