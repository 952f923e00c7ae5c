#!/usr/bin/env python
"""Launches every kernel of liblatentcodec.so once inside a cudaProfilerStart/Stop window, for
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_all python tools/ncu_all_kernels.py
(each section is run once before the window, so allocation and lazy module loading stay outside it).

Sections: config 2 (1024 streams of 16x512 at 8 bits: K2, tables, two-visit table, sort, phase A, B1, B2, size scan,
compaction, decoder v2 + redo pass), 8192 streams (decoder v2 throughput build), 10 bits (register-model decoder),
4 bits (dense small-alphabet decoder), K1 and both dequantisers, the serial encoder / generic decoder, the verbatim-mode
phase B, the stateful coder, ContextModel.update_model."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from image_compression_2_b200 import ContextModel, LatentPipeline, _native, codec, coder  # noqa: E402

small = len(sys.argv) > 1 and sys.argv[1] == "small"  # plain-run check with smaller batches


def lat(B, sigma, seed):
    return (torch.randn(B, 16, 512, generator=torch.Generator().manual_seed(seed)) * sigma).cuda()


def sections():
    out = []
    B2 = 64 if small else 1024
    l8 = lat(B2, 0.14, 1000 + 200000)
    p8 = LatentPipeline(n_symbols=256)
    out.append(("cfg2", lambda: p8.roundtrip_device(l8)))
    B4 = 128 if small else 8192
    l8b = lat(B4, 0.14, 1000 + 400000)
    enc4 = p8.encode(p8.quantize(l8b))  # (encoded outside the window: only the throughput-build decoder is captured)
    out.append(("cfg4_share_decode", lambda: p8.decode(enc4.data, enc4.offsets, enc4.nbits, B4)))
    l10 = lat(64 if small else 1024, 0.14, 31)
    p10 = LatentPipeline(n_symbols=1024, quantizer="affine")
    out.append(("10bit", lambda: p10.roundtrip_device(l10)))
    l4 = lat(64 if small else 1024, 0.4, 32)
    p4 = LatentPipeline(n_symbols=16, quantizer="affine")
    out.append(("4bit", lambda: p4.roundtrip_device(l4)))
    idx8 = p8.quantize(l8)

    def quantisers():
        codec.quantize_affine(l8, 8)                                   # K1, int32 + fp32
        codec.quantize_affine(l8, 8, want_wq=False, idx_dtype=torch.uint8)
        codec.dequantize_affine(idx8, 8)
        codec.dequantize_codebook(idx8, p8.codebook)
        codec.quantize_codebook(l8, p8.codebook.flip(0).contiguous(), sorted_ascending=False)  # full-scan path
    out.append(("quantisers", quantisers))
    sub = idx8[:32].to(torch.int32).contiguous()
    layout = codec.layout_independent(sub.shape)

    def serial():
        enc = codec.encode_batch(sub.reshape(-1), layout, 256, flags=_native.FLAG_ENC_SERIAL)
        codec.decode_batch(enc.data, enc.offsets, enc.nbits, layout, 256, flags=_native.FLAG_DEC_SERIAL)
        codec.decode_batch(enc.data, enc.offsets, enc.nbits, layout, 256, flags=_native.FLAG_DEC_REGISTER_MODEL)
        codec.decode_batch(enc.data, enc.offsets, enc.nbits, layout, 256, flags=_native.FLAG_DEC_GENERIC_SHAPE)
        codec.encode_batch(sub[:, :1, :64].contiguous().reshape(-1), codec.layout_independent((32, 1, 64)), 256, mode="verbatim")
    out.append(("serial_and_variants", serial))
    codes = np.clip(np.round(np.random.default_rng(1).normal(32, 3, (1, 4, 64))), 0, 63).astype(np.int32)

    def stateful():
        cm = ContextModel(64, track_state=True)
        packed, _ = coder.cabac_encode_packed(codes, cm)
        coder.cabac_decode(packed, ContextModel(64, track_state=True), codes.shape)
        cm.update_model((1, 2), 3)
    out.append(("stateful", stateful))
    return out


secs = sections()
for name, fn in secs:  # warm-up outside the profiled window
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, fn in secs:
    fn()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok: %d sections" % len(secs))
