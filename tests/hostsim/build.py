"""TEST INFRASTRUCTURE ONLY -- builds and wraps the CPU SIMT emulation of the coder device code."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_build", "libhostsim.so")
SRCS = [os.path.join(HERE, "hostsim.cpp"), os.path.join(HERE, "cuda_emul.h"),
        os.path.join(ROOT, "image_compression_2_b200", "csrc", "lc_coder.cuh"),
        os.path.join(ROOT, "image_compression_2_b200", "csrc", "lc_common.cuh"),
        os.path.join(ROOT, "image_compression_2_b200", "csrc", "lc_encoder_par.cuh"),
        os.path.join(ROOT, "image_compression_2_b200", "csrc", "lc_decoder_fast.cuh"),
        os.path.join(ROOT, "image_compression_2_b200", "csrc", "lc_decoder_v2.cuh"),
        os.path.join(ROOT, "image_compression_2_b200", "csrc", "lc_decoder_v2_flat.cuh"),
        os.path.join(ROOT, "image_compression_2_b200", "csrc", "lc_decoder_v3.cuh"),
        os.path.join(ROOT, "image_compression_2_b200", "csrc", "lc_decoder_small.cuh"),
        os.path.join(ROOT, "image_compression_2_b200", "csrc", "lc_encoder_sparse.cuh"),
        os.path.join(ROOT, "image_compression_2_b200", "csrc", "lc_encoder_pack.cuh")]
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in SRCS):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               "-Wno-unknown-pragmas", "-I", HERE, "-o", SO, SRCS[0]])
    L = ctypes.CDLL(SO)
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _shape(shape):
    if len(shape) == 4:  # (B streams, imgs, R, C)
        return shape[0], shape[1], shape[2], shape[3], 1
    if len(shape) == 3:  # B independent streams of one image each
        return shape[0], 1, shape[1], shape[2], 1
    if len(shape) == 2:  # B streams, single global context
        return shape[0], 1, 1, shape[1], 0
    raise ValueError(shape)


def encode(codes, n, mode=1, rate=0.05, slot_bytes=None, grid=None):
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    B, imgs, R, C, has_ctx = _shape(codes.shape)
    total = imgs * R * C
    if slot_bytes is None:
        slot_bytes = ((total * 12 + 64) + 3) // 4 * 4
    grid = grid or min(B, 3)
    out = np.zeros((B, slot_bytes), np.uint8)
    nbits = np.zeros(B, np.int32); status = np.zeros(B, np.int32); fault = np.zeros(B, np.int32)
    rc = lib().hostsim_encode(_p(codes, ctypes.c_int), B, imgs, R, C, int(n), ctypes.c_double(rate), int(mode), has_ctx,
                              _p(out, ctypes.c_ubyte), ctypes.c_uint(slot_bytes), _p(nbits, ctypes.c_int),
                              _p(status, ctypes.c_int), _p(fault, ctypes.c_int), grid)
    assert rc == 0, rc
    return out, nbits, status, fault


def decode(streams, n, shape, mode=1, rate=0.05, grid=None, codebook=None, fast=False):
    """streams: list of bytes objects, one per stream; shape: per-batch shape as for encode().
    fast=True: the fast decoder kernel + generic redo pass (repaired mode, 3-D shapes only);
    fast="v2": decoder v2 (decoder warp + updater warps) + generic redo pass."""
    B, imgs, R, C, has_ctx = _shape(shape)
    assert len(streams) == B
    total = imgs * R * C
    offs = np.zeros(B, np.int64)
    nbits = np.zeros(B, np.int32)
    parts, pos = [], 0
    for i, s in enumerate(streams):  # 16-byte aligned stream starts, like the compacted layout
        offs[i] = pos
        nbits[i] = len(s) * 8
        pad = (-len(s)) % 16
        parts.append(bytes(s) + b"\xee" * pad)  # garbage padding: the reader must mask it
        pos += len(s) + pad
    raw = b"".join(parts) + b"\xee" * 16
    blob = np.frombuffer(raw, dtype=np.uint8).copy()
    out = np.full((B, total), -7, np.int32)
    status = np.zeros(B, np.int32); fault = np.zeros(B, np.int32)
    grid = grid or min(B, 3)
    if codebook is not None:
        cb = np.ascontiguousarray(codebook, np.float32)
        deq = np.full((B, total), np.nan, np.float32)
        cbp, deqp = _p(cb, ctypes.c_float), _p(deq, ctypes.c_float)
    else:
        deq, cbp, deqp = None, None, None
    if fast:
        assert mode == 1 and has_ctx
        redone = ctypes.c_int(0)
        if fast in ("v2", "v3"):
            lib().hostsim_set_decoder_version(int(fast[1]))
        fn = (lib().hostsim_decode_v2 if fast in ("v2", "v3") else
              lib().hostsim_decode_small if fast == "small" else lib().hostsim_decode_fast)
        rc = fn(_p(blob, ctypes.c_ubyte), _p(offs, ctypes.c_longlong), _p(nbits, ctypes.c_int), B,
                                       imgs, R, C, int(n), ctypes.c_double(rate), _p(out, ctypes.c_int), cbp, deqp,
                                       _p(status, ctypes.c_int), _p(fault, ctypes.c_int), grid, ctypes.byref(redone))
        assert rc == 0, rc
        decode.last_redone = redone.value
        return out.reshape(shape), status, fault, deq
    rc = lib().hostsim_decode(_p(blob, ctypes.c_ubyte), _p(offs, ctypes.c_longlong), _p(nbits, ctypes.c_int), B, imgs, R, C, int(n),
                              ctypes.c_double(rate), int(mode), has_ctx, _p(out, ctypes.c_int), cbp, deqp,
                              _p(status, ctypes.c_int), _p(fault, ctypes.c_int), grid)
    assert rc == 0, rc
    return out.reshape(shape), status, fault, deq


def encode_par(codes, n, mode=1, rate=0.05, slot_bytes=None, grid=None, nwarps=3):
    """Parallel encoder (phase S on the host, phases A/B emulated). codes: (B,R,C) or (B,imgs,R,C)."""
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    B, imgs, R, C, has_ctx = _shape(codes.shape)
    assert has_ctx
    total = imgs * R * C
    if slot_bytes is None:
        slot_bytes = ((total * 12 + 64) + 3) // 4 * 4
    grid = grid or min(B, 3)
    out = np.zeros((B, slot_bytes), np.uint8)
    nbits = np.zeros(B, np.int32); status = np.zeros(B, np.int32); fault = np.zeros(B, np.int32)
    rc = lib().hostsim_encode_par(_p(codes, ctypes.c_int), B, imgs, R, C, int(n), ctypes.c_double(rate), int(mode),
                                  _p(out, ctypes.c_ubyte), ctypes.c_uint(slot_bytes), _p(nbits, ctypes.c_int),
                                  _p(status, ctypes.c_int), _p(fault, ctypes.c_int), grid, nwarps)
    assert rc == 0, rc
    return out, nbits, status, fault
