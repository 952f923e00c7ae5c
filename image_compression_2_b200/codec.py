"""Tensor-level API over the C ABI (include/latentcodec.h): device tensors in, device tensors out.

PyTorch is used for device memory and streams only; all arithmetic happens in the hand-written
CUDA kernels of liblatentcodec.so.  Every function refuses non-CUDA tensors: there is no CPU path.
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import _native

MODE_VERBATIM, MODE_REPAIRED = 0, 1
MODES = {"verbatim": MODE_VERBATIM, "repaired": MODE_REPAIRED, 0: 0, 1: 1}

STATUS_OK = 0
STATUS_NAMES = {0: "OK", 1: "ENC_BIT_OVERFLOW", 2: "DEC_SYMBOL_OOB", 3: "DEC_ZERO_RANGE", 4: "DEC_NEG_SYMBOL",
                5: "OUT_OVERFLOW", 6: "BAD_SYMBOL", 7: "POOL_OVERFLOW"}


IDX_DTYPES = {torch.int32: 4, torch.int16: 2, torch.uint8: 1}
if hasattr(torch, "uint16"):
    IDX_DTYPES[torch.uint16] = 2


def idx_dtype_for(n_symbols):
    """Narrowest index dtype that holds every symbol of an n-symbol alphabet: uint8 up to 256 symbols, else int16
    (the uint16 bit pattern; alphabets end at 1024 symbols, so the sign bit is never set)."""
    return torch.uint8 if int(n_symbols) <= 256 else torch.int16


class _On:
    """Makes the device the tensors live on current for the duration of a C-ABI call (the library launches on
    the current device, include/latentcodec.h) and hands out that device's current stream."""

    def __init__(self, *tensors):
        devs = {t.device for t in tensors if isinstance(t, torch.Tensor) and t.is_cuda}
        if len(devs) != 1:
            raise RuntimeError("tensors of one call must live on exactly one CUDA device, got %s" % sorted(map(str, devs)))
        self.device = devs.pop()
        self._guard = torch.cuda.device(self.device)

    def __enter__(self):
        self._guard.__enter__()
        return self

    def __exit__(self, *exc):
        return self._guard.__exit__(*exc)

    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream


def _need_cuda(t, name, dtype):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: this package has no CPU path" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def _need_idx(t, name):
    t = _need_cuda(t, name, None)
    if t.dtype not in IDX_DTYPES:
        raise TypeError("%s must be int32, int16/uint16 or uint8, got %s" % (name, t.dtype))
    return t, IDX_DTYPES[t.dtype]


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
# quantisers
# ------------------------------------------------------------------------------------------------

def quantize_affine(w, bits, want_idx=True, want_wq=True, idx_dtype=torch.int32):
    """Quantiser A (stylegan3_hvae_full.py:313-316). Returns (idx | None, wq fp32 | None).  idx_dtype int32 keeps
    the reference's unclamped index; uint8 / int16 hold the index clamped to [0, 2^bits-1] (what the coder takes)."""
    lib = _native.load()
    w = _need_cuda(w, "w", torch.float32)
    idx = torch.empty(w.shape, dtype=idx_dtype, device=w.device) if want_idx else None
    wq = torch.empty_like(w) if want_wq else None
    with _On(w) as on:
        _native.check(lib.lc_quantize_affine_t(w.data_ptr(), w.numel(), int(bits), _ptr(idx), IDX_DTYPES[idx_dtype],
                                               _ptr(wq), on.stream), "lc_quantize_affine")
    return idx, wq


def dequantize_affine(idx, bits):
    """Dequantiser A: idx/(2^bits-1)*2-1 (stylegan3_hvae_full.py:315-316)."""
    lib = _native.load()
    idx, eb = _need_idx(idx, "idx")
    out = torch.empty(idx.shape, dtype=torch.float32, device=idx.device)
    with _On(idx) as on:
        _native.check(lib.lc_dequantize_affine_t(idx.data_ptr(), eb, idx.numel(), int(bits), out.data_ptr(), on.stream),
                      "lc_dequantize_affine")
    return out


def codebook_is_sorted(codebook):
    cb = codebook.detach().float().cpu()
    return bool((cb[1:] >= cb[:-1]).all()) if cb.numel() > 1 else True


def quantize_codebook(z, codebook, want_deq=False, sorted_ascending=None, idx_dtype=torch.int32):
    """Quantiser B (gumbel_softmax_compression.py:97,118): first-minimum argmin over the codebook.
    Returns (idx, codebook[idx] fp32 | None)."""
    lib = _native.load()
    z = _need_cuda(z, "z", torch.float32)
    cb = _need_cuda(codebook.to(z.device), "codebook", torch.float32)
    if sorted_ascending is None:
        sorted_ascending = codebook_is_sorted(cb)
    idx = torch.empty(z.shape, dtype=idx_dtype, device=z.device)
    deq = torch.empty_like(z) if want_deq else None
    with _On(z, cb) as on:
        _native.check(lib.lc_quantize_codebook_t(z.data_ptr(), z.numel(), cb.data_ptr(), cb.numel(),
                                                 1 if sorted_ascending else 0, idx.data_ptr(), IDX_DTYPES[idx_dtype],
                                                 _ptr(deq), on.stream), "lc_quantize_codebook")
    return idx, deq


def dequantize_codebook(idx, codebook):
    """Dequantiser B: codebook[idx] (cabac_compression.py:531)."""
    lib = _native.load()
    idx, eb = _need_idx(idx, "idx")
    cb = _need_cuda(codebook.to(idx.device), "codebook", torch.float32)
    out = torch.empty(idx.shape, dtype=torch.float32, device=idx.device)
    with _On(idx, cb) as on:
        _native.check(lib.lc_dequantize_codebook_t(idx.data_ptr(), eb, idx.numel(), cb.data_ptr(), cb.numel(),
                                                   out.data_ptr(), on.stream), "lc_dequantize_codebook")
    return out


# ------------------------------------------------------------------------------------------------
# coder
# ------------------------------------------------------------------------------------------------

@dataclass
class StreamLayout:
    """How a batch tensor maps onto independent streams."""
    B: int        # independent streams (fresh model each)
    imgs: int     # images per stream sharing coder + model
    R: int
    C: int
    has_ctx: int

    @property
    def total(self):
        return self.imgs * self.R * self.C


def layout_independent(shape):
    """[B,R,C] -> B independent streams (the batch semantics this framework adds)."""
    if len(shape) != 3:
        raise ValueError("expected [B,R,C], got %s" % (tuple(shape),))
    return StreamLayout(int(shape[0]), 1, int(shape[1]), int(shape[2]), 1)


def layout_reference(shape):
    """The reference's own semantics for cabac_encode(data): ONE stream; a 3-D shape (B,R,C) has
    (left,up) contexts per image with a shared model (cabac_compression.py:91-114, 330-337); any
    other rank uses the single global context (:115-117)."""
    shape = tuple(int(s) for s in shape)
    if len(shape) == 3:
        return StreamLayout(1, shape[0], shape[1], shape[2], 1)
    total = int(np.prod(shape)) if len(shape) else 1
    return StreamLayout(1, 1, 1, total, 0)


@dataclass
class EncodedBatch:
    """Device-resident result of encode_batch."""
    data: torch.Tensor      # uint8, compacted streams (16-byte aligned starts)
    offsets: torch.Tensor   # int64 [B+1]
    nbits: torch.Tensor     # int32 [B]
    status: torch.Tensor    # int32 [B]
    fault_index: torch.Tensor  # int32 [B]
    layout: StreamLayout
    n_symbols: int
    mode: int

    def to_host(self):
        """-> (list of packed `bytes` per stream, nbits ndarray, status ndarray, fault ndarray); synchronises."""
        offs = self.offsets.cpu().numpy()
        nbits = self.nbits.cpu().numpy()
        status = self.status.cpu().numpy()
        fault = self.fault_index.cpu().numpy()
        used = int(offs[-1])
        blob = self.data[:used].cpu().numpy()
        out = []
        for b in range(self.layout.B):
            nb = (int(nbits[b]) + 7) // 8 if status[b] == 0 else 0
            out.append(blob[offs[b]:offs[b] + nb].tobytes())
        return out, nbits, status, fault


class CoderWorkspace:
    """Caches the scratch / slot / output buffers for a (device, layout, n) so steady-state calls
    allocate nothing.  A workspace belongs to ONE CUDA stream at a time: two streams (or host threads) coding
    concurrently need a workspace each, or the kernels of one call overwrite the scratch of the other."""

    def __init__(self):
        self._bufs = {}

    def get(self, key, nbytes, device, dtype=torch.uint8):
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes or buf.device != device:
            buf = torch.empty(int(nbytes), dtype=dtype, device=device)
            self._bufs[key] = buf
        return buf


# OR-ed into the `flags` of every encode / decode call: how tests and tools drive one particular kernel over
# inputs that reach the library through the drop-in functions (_native.FLAG_*); results never depend on them
DEFAULT_ENCODE_FLAGS = 0
DEFAULT_DECODE_FLAGS = 0

_default_ws = {}


def default_workspace(device):
    """The workspace of calls that do not bring their own: one per (device, current CUDA stream, host thread), so
    concurrent callers never share scratch."""
    import threading
    key = (str(device), torch.cuda.current_stream(device).cuda_stream, threading.get_ident())
    ws = _default_ws.get(key)
    if ws is None:
        ws = _default_ws[key] = CoderWorkspace()
    return ws


def _check_n(n):
    n = int(n)
    if n < 2 or n > 1024 or (n & (n - 1)):
        raise ValueError("n_symbols must be a power of two in [2,1024] on the CUDA path, got %d" % n)
    return n


def worst_case_slot_bytes(layout):
    """A slot no stream can overflow: the coder emits at most ~80 bits per symbol."""
    return (layout.total * 12 + 256 + 15) // 16 * 16


def encode_batch(idx, layout, n_symbols, mode="repaired", adaptation_rate=0.05, slot_bytes=None, workspace=None,
                 flags=0, reuse_output=False):
    """cabac_encode for layout.B independent streams (cabac_compression.py:315-359).
    idx: int32 / int16 / uint8 CUDA tensor with layout.B * layout.total elements. Returns EncodedBatch (async).
    reuse_output: the result tensors are buffers of `workspace` that its next encode_batch call overwrites (no
    allocation per call)."""
    lib = _native.load()
    idx, eb = _need_idx(idx, "idx")
    n = _check_n(n_symbols)
    if idx.numel() != layout.B * layout.total:
        raise ValueError("idx has %d elements, layout needs %d" % (idx.numel(), layout.B * layout.total))
    dev = idx.device
    ws = workspace or default_workspace(dev)
    B = layout.B
    m = MODES[mode]
    if B == 0:
        z = torch.zeros(0, dtype=torch.int32, device=dev)
        return EncodedBatch(torch.zeros(0, dtype=torch.uint8, device=dev), torch.zeros(1, dtype=torch.int64, device=dev),
                            z, z.clone(), z.clone(), layout, n, m)
    scratch_bytes = lib.lc_coder_scratch_bytes(B, layout.imgs, layout.R, layout.C, n, layout.has_ctx)
    if scratch_bytes < 0:
        raise ValueError("unsupported stream shape for the CUDA coder: %s" % (layout,))
    if slot_bytes is None:
        slot_bytes = lib.lc_encode_slot_bytes(layout.imgs, layout.R, layout.C, n)
    slot_bytes = (int(slot_bytes) + 15) // 16 * 16
    scratch = ws.get(("scratch", dev), scratch_bytes, dev)
    slots = ws.get(("slots", dev), B * slot_bytes, dev)
    if reuse_output:
        out = ws.get(("enc_out", dev), B * slot_bytes, dev)[:B * slot_bytes]
        offsets = ws.get(("enc_offsets", dev), B + 1, dev, torch.int64)[:B + 1]
        meta = ws.get(("enc_meta", dev), 3 * B, dev, torch.int32)
        nbits, status, fault = meta[:B], meta[B:2 * B], meta[2 * B:3 * B]
    else:
        out = torch.empty(B * slot_bytes, dtype=torch.uint8, device=dev)
        offsets = torch.empty(B + 1, dtype=torch.int64, device=dev)
        nbits = torch.empty(B, dtype=torch.int32, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        fault = torch.empty(B, dtype=torch.int32, device=dev)
    with _On(idx) as on:
        rc = lib.lc_encode_batch_t(idx.data_ptr(), eb, B, layout.imgs, layout.R, layout.C, n, float(adaptation_rate), m,
                                   layout.has_ctx, scratch.data_ptr(), scratch.numel(), slots.data_ptr(), slot_bytes,
                                   out.data_ptr(), out.numel(), offsets.data_ptr(), nbits.data_ptr(), status.data_ptr(),
                                   fault.data_ptr(), int(flags) | DEFAULT_ENCODE_FLAGS, on.stream)
    _native.check(rc, "lc_encode_batch")
    return EncodedBatch(out, offsets, nbits, status, fault, layout, n, m)


def encode_batch_checked(idx, layout, n_symbols, **kw):
    """encode_batch, then (synchronising) a second pass with worst-case slots if any stream did not fit its slot:
    skewed streams can need more than the default 1.5*(log2 n + 2) bits per symbol, and the reference never fails
    on them.  Returns (EncodedBatch, (streams, nbits, status, fault) on the host)."""
    enc = encode_batch(idx, layout, n_symbols, **kw)
    host = enc.to_host()
    if (host[2] == 5).any():
        kw = dict(kw, slot_bytes=worst_case_slot_bytes(layout))
        enc = encode_batch(idx, layout, n_symbols, **kw)
        host = enc.to_host()
    return enc, host


def decode_batch(data, offsets, nbits, layout, n_symbols, mode="repaired", adaptation_rate=0.05, codebook=None,
                 workspace=None, deq_out=None, flags=0, idx_dtype=torch.int32, want_idx=True, reuse_output=False):
    """cabac_decode for layout.B independent streams (cabac_compression.py:363-406).
    data uint8 CUDA (stream b = ceil(nbits[b]/8) bytes at offsets[b], offsets multiples of 4, buffer padded to a
    multiple of 4); offsets int64 [>=B]; nbits int32 [B].
    deq_out: optional destination of the dequantised values instead of a fresh CUDA tensor -- fp32, contiguous,
    B*total elements, either on the device or in PINNED host memory (the kernel then writes the rows straight over
    PCIe as it decodes them: no separate device-to-host copy afterwards).
    idx_dtype: int32 (the reference's), int16 or uint8 (n <= 256); want_idx=False skips the index output (needs a
    codebook).  Returns (idx [B,total] | None, deq fp32 [B,total] | None, status int32 [B], fault_index int32 [B])."""
    lib = _native.load()
    data = _need_cuda(data, "data", torch.uint8)
    offsets = _need_cuda(offsets, "offsets", torch.int64)
    nbits = _need_cuda(nbits, "nbits", torch.int32)
    n = _check_n(n_symbols)
    dev = data.device
    ws = workspace or default_workspace(dev)
    B = layout.B
    if not want_idx and codebook is None:
        raise ValueError("want_idx=False needs a codebook (nothing would be produced)")
    if reuse_output:
        idx = ws.get(("dec_idx", dev, idx_dtype), B * layout.total, dev, idx_dtype)[:B * layout.total].view(B, layout.total) \
            if want_idx else None
        meta = ws.get(("dec_meta", dev), 2 * B, dev, torch.int32)
        status, fault = meta[:B], meta[B:2 * B]
    else:
        idx = torch.empty((B, layout.total), dtype=idx_dtype, device=dev) if want_idx else None
        status = torch.empty(B, dtype=torch.int32, device=dev)
        fault = torch.empty(B, dtype=torch.int32, device=dev)
    deq, cb = None, None
    if codebook is not None:
        cb = _need_cuda(codebook.to(dev), "codebook", torch.float32)
        if cb.numel() < n:
            raise ValueError("codebook has %d entries, need %d" % (cb.numel(), n))
        if deq_out is not None:
            if deq_out.dtype != torch.float32 or not deq_out.is_contiguous() or deq_out.numel() != B * layout.total:
                raise ValueError("deq_out must be contiguous fp32 with %d elements" % (B * layout.total))
            if not (deq_out.is_cuda or deq_out.is_pinned()):
                raise RuntimeError("deq_out must be a CUDA tensor or pinned host memory")
            deq = deq_out.view(B, layout.total)
        elif reuse_output:
            deq = ws.get(("dec_deq", dev), B * layout.total, dev, torch.float32)[:B * layout.total].view(B, layout.total)
        else:
            deq = torch.empty((B, layout.total), dtype=torch.float32, device=dev)
    if B == 0:
        return idx, deq, status, fault
    scratch_bytes = lib.lc_coder_scratch_bytes(B, layout.imgs, layout.R, layout.C, n, layout.has_ctx)
    if scratch_bytes < 0:
        raise ValueError("unsupported stream shape for the CUDA coder: %s" % (layout,))
    scratch = ws.get(("scratch", dev), scratch_bytes, dev)
    with _On(data, offsets, nbits, cb, deq if (deq is not None and deq.is_cuda) else None) as on:
        rc = lib.lc_decode_batch_t(data.data_ptr(), offsets.data_ptr(), nbits.data_ptr(), B, layout.imgs, layout.R,
                                   layout.C, n, float(adaptation_rate), MODES[mode], layout.has_ctx, scratch.data_ptr(),
                                   scratch.numel(), _ptr(idx), IDX_DTYPES[idx_dtype], _ptr(cb), _ptr(deq),
                                   status.data_ptr(), fault.data_ptr(), int(flags) | DEFAULT_DECODE_FLAGS, on.stream)
    _native.check(rc, "lc_decode_batch")
    return idx, deq, status, fault


def pack_streams_for_device(streams, device):
    """Host helper: lay out a list of packed `bytes` with 16-byte aligned starts and upload.
    Returns (data uint8, offsets int64 [B+1], nbits int32 [B]) on `device`."""
    B = len(streams)
    offs = np.zeros(B + 1, np.int64)
    nbits = np.zeros(B, np.int32)
    pos = 0
    for i, s in enumerate(streams):
        offs[i] = pos
        nbits[i] = len(s) * 8
        pos += (len(s) + 15) // 16 * 16
    offs[B] = pos
    blob = np.zeros(max(pos, 16), np.uint8)
    for i, s in enumerate(streams):
        blob[offs[i]:offs[i] + len(s)] = np.frombuffer(s, dtype=np.uint8)
    return (torch.from_numpy(blob).to(device), torch.from_numpy(offs).to(device), torch.from_numpy(nbits).to(device))
