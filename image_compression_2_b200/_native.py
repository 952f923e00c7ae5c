"""ctypes binding of liblatentcodec.so (C ABI: include/latentcodec.h).

There is no CPU fallback: if the CUDA library cannot be loaded the import of the ops fails
loudly, and every op refuses tensors that are not on a CUDA device.
"""
import ctypes
import os

from . import build as _build

_lib = None

_vp, _i32, _i64, _dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double

SIGNATURES = {
    "lc_version": (ctypes.c_int, []),
    "lc_debug_launch_count": (_i64, [_i32]),
    "lc_quantize_affine_t": (ctypes.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp]),
    "lc_dequantize_affine_t": (ctypes.c_int, [_vp, _i32, _i64, _i32, _vp, _vp]),
    "lc_quantize_codebook_t": (ctypes.c_int, [_vp, _i64, _vp, _i32, _i32, _vp, _i32, _vp, _vp]),
    "lc_dequantize_codebook_t": (ctypes.c_int, [_vp, _i32, _i64, _vp, _i32, _vp, _vp]),
    "lc_encode_batch_t": (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _dbl, _i32, _i32, _vp, _i64, _vp, _i64,
                                         _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp]),
    "lc_decode_batch_t": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _dbl, _i32, _i32, _vp, _i64,
                                         _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "lc_quantize_affine": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "lc_dequantize_affine": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "lc_quantize_codebook": (ctypes.c_int, [_vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp]),
    "lc_dequantize_codebook": (ctypes.c_int, [_vp, _i64, _vp, _i32, _vp, _vp]),
    "lc_coder_scratch_bytes": (_i64, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "lc_encode_slot_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "lc_coder_grid": (ctypes.c_int, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "lc_encode_batch": (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _dbl, _i32, _i32, _vp, _i64, _vp, _i64,
                                       _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "lc_model_update": (ctypes.c_int, [_vp, _i32, _i32, _dbl, _vp]),
    "lc_stateful_table_bytes": (_i64, [_i32, _i32]),
    "lc_stateful_table_offset": (_i64, [_i32, _i32, _i32]),
    "lc_stateful_encode": (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _dbl, _i32, _i32, _vp, _i64, _vp, _i64, _vp, _vp,
                                          _vp, _vp]),
    "lc_stateful_decode": (ctypes.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _dbl, _i32, _i32, _vp, _i64, _vp, _vp, _vp,
                                          _vp]),
    "lc_decode_batch": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _dbl, _i32, _i32, _vp, _i64,
                                       _vp, _vp, _vp, _vp, _vp, _vp]),
}


ABI_VERSION = 2

# `flags` of the _t entry points (include/latentcodec.h)
FLAG_DEC_LATENCY_BUILD, FLAG_DEC_THROUGHPUT_BUILD, FLAG_DEC_GENERIC_SHAPE = 1, 2, 4
FLAG_DEC_REGISTER_MODEL, FLAG_DEC_SERIAL, FLAG_ENC_SERIAL, FLAG_DEC_NO_SMALL = 8, 16, 32, 64
FLAG_ENC_SORT_V1 = 128


class NativeLibraryError(RuntimeError):
    pass


def library_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load (building first if needed and possible) the CUDA library. Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    override = os.environ.get("LATENTCODEC_LIB")  # debug builds (tools/dec_profile.py); same ABI
    if override:
        path, build_if_missing = override, False
    if build_if_missing and _build.needs_build():
        if _build.nvcc_path() is not None:
            _build.build_library()
        elif not os.path.exists(path):
            raise NativeLibraryError("liblatentcodec.so is not built and nvcc is not available; "
                                     "run `python -m image_compression_2_b200.build`")
    if not os.path.exists(path):
        raise NativeLibraryError("liblatentcodec.so missing at %s (no CPU fallback exists)" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.lc_version() != ABI_VERSION:
        raise NativeLibraryError("liblatentcodec.so ABI version %d, expected %d" % (lib.lc_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what):
    if rc == 0:
        return
    if rc == -22:
        raise ValueError("%s: unsupported or invalid arguments (EINVAL)" % what)
    if rc == -12:
        raise MemoryError("%s: scratch buffer too small" % what)
    if rc <= -1000:
        raise RuntimeError("%s: CUDA error %d" % (what, -1000 - rc))
    raise RuntimeError("%s failed with code %d" % (what, rc))
