"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF.

Run in the build container (where /root/reference exists):  python oracle/make_golden.py
Every array is produced by the unmodified reference code imported through oracle/ref_loader.py
(torch / numpy versions are recorded in the fixture); the C oracle is not involved.  The
fixtures are what pins both the C oracle and the CUDA path on the GPU box, where the reference
tree does not exist.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def ref_quant_a(w, bits):
    """StyleGAN3Compressor.compress with the encoder stubbed to return `w` as the means
    (stylegan3_hvae_full.py:295-318)."""
    R.load()
    import stylegan3_hvae_full as S

    class Enc(torch.nn.Module):
        def forward(self, x):
            return x, x, None

    class Gen(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

    comp = S.StyleGAN3Compressor(Enc(), Gen())
    with torch.no_grad():
        return comp.compress(torch.from_numpy(w), quantization_bits=bits, deterministic=True).numpy()


def ref_quant_a_idx(w, bits):
    """The integer the reference never materialises: torch.round(((w+1)*0.5)*scale) (:313-315)."""
    t = torch.from_numpy(w)
    scale = (2 ** bits) - 1
    return torch.round(((t + 1) * 0.5) * scale).numpy()


def ref_quant_b(z, n):
    """GumbelSoftmaxDiscretization.forward -> encoding_indices (gumbel_softmax_compression.py:93-118)
    and codebook[idx] (cabac_compression.py:531)."""
    R.load()
    import gumbel_softmax_compression as G
    torch.manual_seed(0)
    d = G.GumbelSoftmaxDiscretization(latent_dim=z.shape[-1], n_embeddings=n)
    d.eval()
    with torch.no_grad():
        _, _, idx = d(torch.from_numpy(z), hard=True)
        deq = d.codebook[idx]
    return idx.numpy().astype(np.int32).reshape(z.shape), deq.numpy().reshape(z.shape), d.codebook.numpy().copy()


def coder_case(codes, n, mode):
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    bits, err = R.ref_encode_bits(codes, n, mode)
    rec = {"codes": codes, "n": np.int32(n)}
    if err is not None:
        rec["enc_error"] = np.array([err[0]])
        rec["enc_fault_index"] = np.int64(err[1])
        return rec
    packed = np.frombuffer(R.pack_bits(bits), dtype=np.uint8)
    rec["nbits"] = np.int64(len(bits))
    rec["packed"] = packed
    dec, derr = R.ref_decode(packed.tobytes(), n, codes.shape, mode)
    if derr is not None:
        rec["dec_error"] = np.array([derr[0]])
        rec["dec_fault_index"] = np.int64(derr[1])
        rec["decoded"] = dec
    else:
        rec["decoded"] = dec
    return rec


def save(name, **arrs):
    arrs["_versions"] = np.array(["numpy " + np.__version__, "torch " + torch.__version__])
    np.savez_compressed(os.path.join(OUT, name), **arrs)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrs.items() if not k.startswith("_")})


def flat(prefix, rec):
    return {prefix + "__" + k: v for k, v in rec.items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)

    # ---- known-answer vectors from SURVEY.md section 8c
    kat = {}
    k1 = np.array([0, 3, 2, 1] * 4, dtype=np.int32).reshape(1, 2, 8)
    k2 = np.random.default_rng(0).integers(0, 256, (1, 2, 16)).astype(np.int32)
    z3 = (torch.randn(1, 16, 512, generator=torch.Generator().manual_seed(1234)) * 0.14).numpy()
    idx3, deq3, cb256 = ref_quant_b(z3, 256)
    for mode in ("verbatim", "repaired"):
        kat.update(flat("kat1_" + mode, coder_case(k1, 4, mode)))
        kat.update(flat("kat2_" + mode, coder_case(k2, 256, mode)))
    kat.update(flat("kat3_repaired", coder_case(idx3, 256, "repaired")))
    kat.update(flat("kat3_verbatim", coder_case(idx3, 256, "verbatim")))
    kat["kat3__z"] = z3
    kat["kat3__deq"] = deq3
    assert hashlib.sha256(idx3.tobytes()).hexdigest()[:16] == "910821543ca43087"
    assert hashlib.sha256(kat["kat3_repaired__packed"].tobytes()).hexdigest()[:16] == "3967e5e858e1138c"
    save("kat.npz", **kat)

    # ---- config 1: the real random-init HVAE_VGG_Encoder sample (means captured once)
    R.load()
    import contextlib
    import io
    import stylegan3_hvae_full as S
    torch.manual_seed(0)
    enc = S.HVAE_VGG_Encoder(img_resolution=1024)
    enc.eval()
    x = torch.randn(1, 3, 256, 256).clamp(-1, 1)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        means = enc(x)[1].numpy().astype(np.float32)
    cfg1 = {"means": means}
    for bits in (4, 8, 10):
        cfg1["a_wq_%d" % bits] = ref_quant_a(means, bits)
        cfg1["a_idx_%d" % bits] = ref_quant_a_idx(means, bits)
    idx, deq, cb = ref_quant_b(means, 256)
    cfg1["b_idx_256"], cfg1["b_deq_256"], cfg1["codebook_256"] = idx, deq, cb
    cfg1.update(flat("coder_repaired", coder_case(idx, 256, "repaired")))
    cfg1.update(flat("coder_verbatim", coder_case(idx, 256, "verbatim")))
    save("config1.npz", **cfg1)

    # ---- quantiser fixtures: specials, ties, out-of-range, random
    q = {}
    specials = np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 1.5, -1.5, 3.0, -3.0, 1e-8, -1e-8, 0.99999994, -0.99999994,
                         1e10, -1e10, np.inf, -np.inf, np.nan, 2.0 / 255, 1.0 / 255, 3.0 / 255,
                         np.float32(1.0 / 15), np.float32(-7.0 / 15)], dtype=np.float32)
    w = np.concatenate([specials,
                        (rng.standard_normal(6000) * 0.4).astype(np.float32),
                        (rng.random(3000) * 2 - 1).astype(np.float32),
                        (rng.standard_normal(3000) * 0.14).astype(np.float32)])
    # exact half-way points of quantiser A at each bit depth: (2j+1)/(2*scale)*2-1
    for bits in (4, 8, 10):
        sc = 2 ** bits - 1
        j = np.arange(0, min(sc, 64))
        w = np.concatenate([w, ((2 * j + 1) / (2.0 * sc) * 2 - 1).astype(np.float32)])
    pad = (-len(w)) % 512
    w = np.concatenate([w, np.zeros(pad, np.float32)]).reshape(1, -1, 512)
    q["w"] = w
    for bits in (4, 6, 8, 10):
        q["a_wq_%d" % bits] = ref_quant_a(w, bits)
        q["a_idx_%d" % bits] = ref_quant_a_idx(w, bits)
    for n in (16, 64, 256, 1024):
        cb = torch.linspace(-1, 1, n).float().numpy()
        mids = ((cb[:-1].astype(np.float64) + cb[1:].astype(np.float64)) / 2).astype(np.float32)
        zz = np.concatenate([w.ravel(), mids, np.nextafter(mids, np.float32(2)), np.nextafter(mids, np.float32(-2)), cb])
        zz = np.concatenate([zz, np.zeros((-len(zz)) % 512, np.float32)]).reshape(1, -1, 512)
        idx, deq, cbo = ref_quant_b(zz, n)
        q["b_z_%d" % n], q["b_idx_%d" % n], q["b_deq_%d" % n], q["codebook_%d" % n] = zz, idx, deq, cbo
    save("quantizers.npz", **q)

    # ---- coder fixtures: full 16x512 streams (cfg 2/3 style) and decoder-hazard streams
    full = {}
    specs = [("b4_wide", 16, 0.4, 11), ("b4_enc_like", 16, 0.14, 12), ("b8_enc_like", 256, 0.14, 13),
             ("b8_wide", 256, 0.4, 14), ("b10_enc_like", 1024, 0.14, 15), ("b6_enc_like", 64, 0.14, 16)]
    for name, n, sig, seed in specs:
        z = (torch.randn(1, 16, 512, generator=torch.Generator().manual_seed(seed)) * sig).numpy()
        idx, _, _ = ref_quant_b(z, n)
        full.update(flat(name, coder_case(idx, n, "repaired")))
    u = rng.integers(0, 256, (1, 16, 512)).astype(np.int32)
    full.update(flat("b8_uniform", coder_case(u, 256, "repaired")))
    const = np.full((1, 16, 512), 7, np.int32)
    full.update(flat("b4_const", coder_case(const, 16, "repaired")))
    save("coder_full.npz", **full)

    # ---- many small streams, both modes, incl. ragged shapes, B>1 shared-model streams,
    #      two-valued hazard streams (H1/H2), and the non-3-D single-context fallback
    small = {}
    cid = 0
    for t in range(120):
        n = int(rng.choice([2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]))
        shape = (int(rng.integers(1, 4)), int(rng.integers(1, 6)), int(rng.integers(1, 48)))
        kind = int(rng.integers(0, 4))
        if kind == 0:
            c = rng.integers(0, n, shape)
        elif kind == 1:
            c = np.clip(np.round(rng.normal(n / 2, max(1.0, n / 16), shape)), 0, n - 1)
        elif kind == 2:
            c = rng.choice([0, n - 1], shape)
        else:
            c = np.clip(np.round(rng.normal(n / 2, 0.7, shape)), 0, n - 1)
        for mode in ("repaired", "verbatim"):
            small.update(flat("s%03d_%s" % (cid, mode), coder_case(c.astype(np.int32), n, mode)))
        cid += 1
    for t in range(24):  # H1/H2 hazards: long two-valued rows
        c = rng.choice([0, 255], (1, 1, 512)).astype(np.int32)
        small.update(flat("h%03d_repaired" % t, coder_case(c, 256, "repaired")))
    for t in range(6):  # non-3-D shapes: one global context, many distinct symbols per context
        n = [16, 64, 256][t % 3]
        c = rng.integers(0, n, (int(rng.integers(100, 600)),)).astype(np.int32)
        small.update(flat("g%03d_repaired" % t, coder_case(c, n, "repaired")))
    save("coder_small.npz", **small)


if __name__ == "__main__":
    main()
