"""Bit-rate statistics over a coded batch.

The reference reports sizes per image inside CABACCompressor.compress (cabac_compression.py:474-489:
`orig_size = size * log2(n) / 8`, `comp_size = len(encoded_bytes)`, ratio) and per method in
compare_compression_methods (:800-881, "hvae" = fixed-length codes vs "cabac" = arithmetic-coded).  This is the
same accounting for a whole batch, computed from the per-stream bit counts the encoder returns; the PNG/JPEG
legs of the reference harness need image files and the generator and stay out of scope.

Note (defect D1): the reference's `comp_size` counts emitted BITS as bytes.  `reference_comp_size`
reproduces that figure; `packed_bytes` is the real size.
"""
import math

import numpy as np
import torch


def bitrate_stats(nbits, symbols_per_stream, n_symbols, status=None, hist_bins=16):
    """nbits: int tensor/array [B] of coded bits per stream (EncodedBatch.nbits); streams whose status is non-zero
    are left out.  Returns a dict of plain Python numbers and lists."""
    nb = nbits.detach().to("cpu", torch.float64).numpy() if isinstance(nbits, torch.Tensor) else np.asarray(nbits, np.float64)
    if status is not None:
        st = status.detach().cpu().numpy() if isinstance(status, torch.Tensor) else np.asarray(status)
        nb = nb[st == 0]
    if nb.size == 0:
        raise ValueError("no successfully coded stream")
    raw_bits = float(symbols_per_stream) * math.log2(n_symbols)  # fixed-length codes ("HVAE standard")
    packed = np.ceil(nb / 8.0)
    hist, edges = np.histogram(nb / symbols_per_stream, bins=hist_bins)
    return {
        "streams": int(nb.size),
        "symbols_per_stream": int(symbols_per_stream),
        "n_symbols": int(n_symbols),
        "raw_bits_per_stream": raw_bits,
        "coded_bits_per_stream": {"mean": float(nb.mean()), "min": float(nb.min()), "max": float(nb.max()),
                                  "p05": float(np.percentile(nb, 5)), "p50": float(np.percentile(nb, 50)),
                                  "p95": float(np.percentile(nb, 95))},
        "coded_bits_per_symbol": float(nb.mean() / symbols_per_stream),
        "packed_bytes": {"mean": float(packed.mean()), "total": int(packed.sum())},
        "orig_size": raw_bits / 8.0,                      # the reference's metadata['orig_size'] per stream
        "reference_comp_size": float(nb.mean()),          # ... and its 'comp_size' (bits counted as bytes, D1)
        "compression_ratio_reference": float(raw_bits / 8.0 / nb.mean()),
        "cabac_vs_raw": float(raw_bits / nb.mean()),      # > 1: the arithmetic coder beats fixed-length codes
        "histogram_bits_per_symbol": {"counts": hist.tolist(), "edges": [float(e) for e in edges]},
    }


def encoded_batch_stats(enc, hist_bins=16):
    """bitrate_stats of a codec.EncodedBatch (device tensors; one small device->host copy)."""
    return bitrate_stats(enc.nbits, enc.layout.total, enc.n_symbols, status=enc.status, hist_bins=hist_bins)
