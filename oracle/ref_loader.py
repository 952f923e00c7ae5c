"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference implementation.

Imports /root/reference/{cabac_compression,stylegan3_hvae_full,gumbel_softmax_compression}.py
with empty stub modules for the un-vendored third-party imports (torch_utils, dnnlib, lpips;
stylegan3_hvae_full.py:9-13,24) so the hot-path code (cabac_compression.py:60-406,
stylegan3_hvae_full.py:313-316, gumbel_softmax_compression.py:49-52,93-118) can be executed
here to (a) validate the C restatement in latent_oracle.c and (b) generate the golden vectors
committed under tests/golden/ (see make_golden.py).

/root/reference only exists in the build container; nothing that runs on the GPU box imports
this module.  Nothing in the product package imports anything under oracle/.

Two coder modes:
  verbatim : the reference file exactly as shipped.
  repaired : ArithmeticCoder subclass changing three tokens --
             `| self.full_range` -> `| self.half_range` in _handle_underflow
             (cabac_compression.py:210) and in decode_symbol's underflow loop (:308), and a
             32-bit mask on the decoder's code_value update (:309).  The encoder's bit list is
             packed MSB-first (np.packbits) before it is handed to the decoder (defects D1/D2).
A "stream" is one array coded with a FRESH ContextModel (defect D5).
"""
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("IC2_REFERENCE_DIR", "/root/reference")

_ref = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "cabac_compression.py"))


def load():
    """Import the reference modules (cached). Returns the cabac_compression module."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
    for name in ("torch_utils", "torch_utils.misc", "dnnlib", "lpips"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["torch_utils"].misc = sys.modules["torch_utils.misc"]
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import cabac_compression as cc  # noqa: E402

    cc.tqdm = lambda it, **k: it
    cc._VerbatimCoder = cc.ArithmeticCoder

    class RepairedCoder(cc._VerbatimCoder):
        def _handle_underflow(self):
            while ((self.low & self.quarter_range) != 0) and ((self.high & self.quarter_range) == 0):
                self.outstanding_bytes += 1
                self.low = (self.low << 1) & (self.half_range - 1)
                self.high = ((self.high << 1) & (self.half_range - 1)) | self.half_range | 1

        def decode_symbol(self, cumulative_probs):
            range_size = self.high - self.low + 1
            scaled_value = ((self.code_value - self.low + 1) * 1.0 / range_size) - 1e-10
            symbol = np.searchsorted(cumulative_probs, scaled_value) - 1
            self.high = self.low + int(range_size * cumulative_probs[symbol + 1] - 1)
            self.low = self.low + int(range_size * cumulative_probs[symbol])
            while (self.high & self.half_range) == (self.low & self.half_range):
                self.low = (self.low << 1) & (self.full_range - 1)
                self.high = ((self.high << 1) & (self.full_range - 1)) | 1
                self.code_value = ((self.code_value << 1) & (self.full_range - 1)) | self._read_bit()
            while ((self.low & self.quarter_range) != 0) and ((self.high & self.quarter_range) == 0):
                self.low = (self.low << 1) & (self.half_range - 1)
                self.high = ((self.high << 1) & (self.half_range - 1)) | self.half_range | 1
                self.code_value = (((self.code_value ^ self.quarter_range) << 1)
                                   & (self.full_range - 1)) | self._read_bit()
            return symbol

    cc._RepairedCoder = RepairedCoder
    _ref = cc
    return cc


def set_mode(mode):
    cc = load()
    if mode == "verbatim":
        cc.ArithmeticCoder = cc._VerbatimCoder
    elif mode == "repaired":
        cc.ArithmeticCoder = cc._RepairedCoder
    else:
        raise ValueError(mode)
    return cc


def ref_encode_bits(codes, n_symbols, mode="repaired"):
    """Run the reference encoder on one stream with a fresh model.

    Returns (bits uint8[ nbits ] of 0/1, error) where error is None or (exception class name,
    number of symbols fully encoded before the exception).
    """
    cc = set_mode(mode)
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    cm = cc.ContextModel(n_symbols=n_symbols)
    # count completed symbols through update_model calls
    done = [0]
    orig_update = cm.update_model

    def counting_update(ctx, sym):
        orig_update(ctx, sym)
        done[0] += 1

    cm.update_model = counting_update
    try:
        out = cc.cabac_encode(codes, cm)
    except (ValueError, IndexError, ZeroDivisionError, OverflowError) as e:
        return None, (type(e).__name__, done[0])
    return np.frombuffer(out, dtype=np.uint8).copy(), None


def ref_decode(packed, n_symbols, shape, mode="repaired"):
    """Run the reference decoder on packed bytes with a fresh model.

    Returns (symbols int32[shape] (entries after a fault are 0), error) with error None or
    (exception class name, number of symbols decoded before the exception).
    """
    cc = set_mode(mode)
    cm = cc.ContextModel(n_symbols=n_symbols)
    captured = {}
    orig_zeros = np.zeros
    done = [0]
    orig_update = cm.update_model

    def counting_update(ctx, sym):
        orig_update(ctx, sym)
        done[0] += 1

    cm.update_model = counting_update

    # capture decoded_flat so partial output survives an exception
    def zeros_spy(*a, **k):
        arr = orig_zeros(*a, **k)
        if k.get("dtype", None) is np.int32 or (len(a) > 1 and a[1] is np.int32):
            captured["flat"] = arr
        return arr

    cc.np.zeros = zeros_spy
    try:
        try:
            out = cc.cabac_decode(bytes(packed), cm, tuple(shape))
        finally:
            cc.np.zeros = orig_zeros
    except (ValueError, IndexError, ZeroDivisionError, OverflowError) as e:
        flat = captured.get("flat")
        part = flat.reshape(shape).copy() if flat is not None else None
        return part, (type(e).__name__, done[0])
    return np.asarray(out, dtype=np.int32), None


def pack_bits(bits):
    """MSB-first packing of the reference encoder's one-byte-per-bit output."""
    return np.packbits(np.asarray(bits, dtype=np.uint8)).tobytes()
