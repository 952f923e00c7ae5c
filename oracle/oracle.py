"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around the C oracle (oracle/latent_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  The product package (image_compression_2_b200) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblatent_oracle.so")

MODE_VERBATIM, MODE_REPAIRED = 0, 1
MODES = {"verbatim": MODE_VERBATIM, "repaired": MODE_REPAIRED}
OK, ENC_BIT_OVERFLOW, DEC_SYMBOL_OOB, DEC_ZERO_RANGE, DEC_NEG_SYMBOL, OUT_OVERFLOW, BAD_SYMBOL = range(7)

_lib = None


def build(force=False):
    src = os.path.join(_HERE, "latent_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i32p, u8p, f32p, i64p = (ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint8),
                                 ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64))
        L.orc_quantize_affine.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, i32p, f32p]
        L.orc_quantize_affine.restype = None
        L.orc_dequantize_affine.argtypes = [i32p, ctypes.c_int64, ctypes.c_int, f32p]
        L.orc_dequantize_affine.restype = None
        L.orc_quantize_codebook.argtypes = [f32p, ctypes.c_int64, f32p, ctypes.c_int, i32p]
        L.orc_quantize_codebook.restype = None
        L.orc_dequantize_codebook.argtypes = [i32p, ctypes.c_int64, f32p, f32p]
        L.orc_dequantize_codebook.restype = None
        L.orc_np_sum.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int64]
        L.orc_np_sum.restype = ctypes.c_double
        L.orc_encode_stream.argtypes = [i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_double, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int64,
                                        i64p, i64p, i64p]
        L.orc_encode_stream.restype = ctypes.c_int
        L.orc_decode_stream.argtypes = [u8p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, i32p, i64p]
        L.orc_decode_stream.restype = ctypes.c_int
        f64p = ctypes.POINTER(ctypes.c_double)
        model_io = [ctypes.c_int64, i32p, i32p, i64p, f64p, ctypes.c_int64, i64p, i32p, i32p, i64p, f64p]
        L.orc_encode_stream_model.argtypes = [i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_double, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int64,
                                              i64p, i64p] + model_io
        L.orc_encode_stream_model.restype = ctypes.c_int
        L.orc_decode_stream_model.argtypes = [u8p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, i32p,
                                              i64p] + model_io
        L.orc_decode_stream_model.restype = ctypes.c_int
        L.orc_pack_bits.argtypes = [u8p, ctypes.c_int64, u8p]
        L.orc_pack_bits.restype = None
        L.orc_roundtrip_stream.argtypes = [i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_double, ctypes.c_int, u8p, ctypes.c_int64, u8p, i32p, i64p]
        L.orc_roundtrip_stream.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def quantize_affine(w, bits):
    """-> (idx int32, dequantised fp32), both shaped like w."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    idx = np.empty(w.shape, np.int32)
    wq = np.empty(w.shape, np.float32)
    lib().orc_quantize_affine(_p(w, ctypes.c_float), w.size, int(bits), _p(idx, ctypes.c_int32), _p(wq, ctypes.c_float))
    return idx, wq


def dequantize_affine(idx, bits):
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.empty(idx.shape, np.float32)
    lib().orc_dequantize_affine(_p(idx, ctypes.c_int32), idx.size, int(bits), _p(out, ctypes.c_float))
    return out


def quantize_codebook(z, codebook):
    z = np.ascontiguousarray(z, dtype=np.float32)
    cb = np.ascontiguousarray(codebook, dtype=np.float32)
    idx = np.empty(z.shape, np.int32)
    lib().orc_quantize_codebook(_p(z, ctypes.c_float), z.size, _p(cb, ctypes.c_float), cb.size, _p(idx, ctypes.c_int32))
    return idx


def dequantize_codebook(idx, codebook):
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    cb = np.ascontiguousarray(codebook, dtype=np.float32)
    out = np.empty(idx.shape, np.float32)
    lib().orc_dequantize_codebook(_p(idx, ctypes.c_int32), idx.size, _p(cb, ctypes.c_float), _p(out, ctypes.c_float))
    return out


def np_sum(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().orc_np_sum(_p(a, ctypes.c_double), a.size)


def _shape3(shape):
    if len(shape) == 3:
        return int(shape[0]), int(shape[1]), int(shape[2]), 1
    total = int(np.prod(shape)) if len(shape) else 1
    return 1, 1, total, 0  # non-3-D: single global context (cabac_compression.py:115-117)


def encode_stream(codes, n_symbols, mode="repaired", rate=0.05, bits_per_symbol_cap=80):
    """One stream, fresh model. -> dict(status, fault_index, nbits, bits (0/1 uint8), packed bytes, n_contexts)."""
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    B, R, C, has_ctx = _shape3(codes.shape)
    cap = codes.size * bits_per_symbol_cap + 64
    bits = np.empty(cap, np.uint8)
    nbits, fault, nctx = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    st = lib().orc_encode_stream(_p(codes, ctypes.c_int32), B, R, C, int(n_symbols), float(rate), MODES[mode],
                                 has_ctx, _p(bits, ctypes.c_uint8), cap, ctypes.byref(nbits),
                                 ctypes.byref(fault), ctypes.byref(nctx))
    bits = bits[:nbits.value].copy()
    return dict(status=st, fault_index=fault.value, nbits=nbits.value, bits=bits,
                packed=np.packbits(bits).tobytes() if st == OK else b"", n_contexts=nctx.value)


def decode_stream(packed, n_symbols, shape, mode="repaired", rate=0.05):
    """One stream, fresh model. -> dict(status, fault_index, symbols int32[shape])."""
    buf = np.frombuffer(bytes(packed), dtype=np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)
        nbytes = 0
    else:
        nbytes = buf.size
    B, R, C, has_ctx = _shape3(tuple(shape))
    out = np.zeros(B * R * C, np.int32)
    fault = ctypes.c_int64(0)
    st = lib().orc_decode_stream(_p(buf, ctypes.c_uint8), nbytes, B, R, C, int(n_symbols), float(rate), MODES[mode],
                                 has_ctx, _p(out, ctypes.c_int32), ctypes.byref(fault))
    return dict(status=st, fault_index=fault.value, symbols=out.reshape(shape))


def roundtrip_stream(codes, n_symbols, mode="repaired", rate=0.05):
    """encode+decode one (B,R,C) stream; returns (status, nbits). Used by the CPU baseline timer."""
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    B, R, C, _ = _shape3(codes.shape)
    cap = codes.size * 80 + 64
    bits = np.empty(cap, np.uint8)
    packed = np.empty(cap // 8 + 8, np.uint8)
    dec = np.empty(codes.size, np.int32)
    nbits = ctypes.c_int64(0)
    st = lib().orc_roundtrip_stream(_p(codes, ctypes.c_int32), B, R, C, int(n_symbols), float(rate), MODES[mode],
                                    _p(bits, ctypes.c_uint8), cap, _p(packed, ctypes.c_uint8),
                                    _p(dec, ctypes.c_int32), ctypes.byref(nbits))
    return st, nbits.value


# ---- stateful model variants (the reference mutates one ContextModel across calls, defect D5) ---------------
# A model is a dict {(left, up): (float64[n] vector, count)}; the global context `()` of non-3-D data is (-2, -2).

def _model_arrays(model, n):
    keys = list(model.keys())
    left = np.array([k[0] for k in keys], np.int32).reshape(-1)
    up = np.array([k[1] for k in keys], np.int32).reshape(-1)
    counts = np.array([int(model[k][1]) for k in keys], np.int64).reshape(-1)
    vecs = np.ascontiguousarray(np.stack([np.asarray(model[k][0], np.float64) for k in keys]) if keys
                                else np.zeros((0, n)), np.float64)
    pad = lambda a: a if a.size else np.zeros(1, a.dtype)  # noqa: E731
    return len(keys), pad(left), pad(up), pad(counts), (vecs if vecs.size else np.zeros((1, n)))


def _model_out(total, nin, n):
    cap = total + nin + 1
    return (cap, ctypes.c_int64(0), np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int64),
            np.zeros((cap, n), np.float64))


def _model_dict(nout, left, up, counts, vecs):
    return {(int(left[i]), int(up[i])): (vecs[i].copy(), int(counts[i])) for i in range(nout)}


def encode_stream_model(codes, n_symbols, model, mode="repaired", rate=0.05):
    """encode_stream starting from `model`; returns (result dict, model after the call)."""
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    B, R, C, has_ctx = _shape3(codes.shape)
    cap_bits = codes.size * 80 + 64
    bits = np.empty(cap_bits, np.uint8)
    nbits, fault = ctypes.c_int64(0), ctypes.c_int64(0)
    nin, il, iu, ic, iv = _model_arrays(model, n_symbols)
    cap, nout, ol, ou, oc, ov = _model_out(codes.size, nin, n_symbols)
    f64 = ctypes.c_double
    st = lib().orc_encode_stream_model(_p(codes, ctypes.c_int32), B, R, C, int(n_symbols), float(rate), MODES[mode],
                                       has_ctx, _p(bits, ctypes.c_uint8), cap_bits, ctypes.byref(nbits),
                                       ctypes.byref(fault), nin, _p(il, ctypes.c_int32), _p(iu, ctypes.c_int32),
                                       _p(ic, ctypes.c_int64), _p(iv, f64), cap, ctypes.byref(nout),
                                       _p(ol, ctypes.c_int32), _p(ou, ctypes.c_int32), _p(oc, ctypes.c_int64), _p(ov, f64))
    bits = bits[:nbits.value].copy()
    res = dict(status=st, fault_index=fault.value, nbits=nbits.value, bits=bits,
               packed=np.packbits(bits).tobytes() if st == OK else b"")
    return res, _model_dict(nout.value, ol, ou, oc, ov)


def decode_stream_model(packed, n_symbols, shape, model, mode="repaired", rate=0.05):
    """decode_stream starting from `model`; returns (result dict, model after the call)."""
    buf = np.frombuffer(bytes(packed), dtype=np.uint8)
    nbytes = buf.size
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)
    B, R, C, has_ctx = _shape3(tuple(shape))
    total = B * R * C
    out = np.zeros(total, np.int32)
    fault = ctypes.c_int64(0)
    nin, il, iu, ic, iv = _model_arrays(model, n_symbols)
    cap, nout, ol, ou, oc, ov = _model_out(total, nin, n_symbols)
    f64 = ctypes.c_double
    st = lib().orc_decode_stream_model(_p(buf, ctypes.c_uint8), nbytes, B, R, C, int(n_symbols), float(rate), MODES[mode],
                                       has_ctx, _p(out, ctypes.c_int32), ctypes.byref(fault), nin,
                                       _p(il, ctypes.c_int32), _p(iu, ctypes.c_int32), _p(ic, ctypes.c_int64), _p(iv, f64),
                                       cap, ctypes.byref(nout), _p(ol, ctypes.c_int32), _p(ou, ctypes.c_int32),
                                       _p(oc, ctypes.c_int64), _p(ov, f64))
    return dict(status=st, fault_index=fault.value, symbols=out.reshape(shape)), _model_dict(nout.value, ol, ou, oc, ov)
