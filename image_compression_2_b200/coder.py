"""Drop-in host surface for the reference's coder functions, backed by the CUDA kernels.

Mirrors /root/reference/cabac_compression.py: ContextModel (:60-162, constructor and state
containers only -- the model itself lives on the GPU), cabac_encode (:315-359), cabac_decode
(:363-406).  Same argument meaning, same return types, same exception classes.

Deviations, all documented in DESIGN.md:
  * only a FRESH (empty) ContextModel is accepted (the reference mutates one shared model across
    calls, defect D5); a non-empty model raises NotImplementedError -- there is no CPU fallback;
  * the model object is not filled in after the call;
  * DEFAULT_MODE is "repaired": the file as shipped ("verbatim") raises ValueError within a few
    hundred symbols on any realistic stream (defect D3).  mode="verbatim" reproduces that.
"""
import numpy as np
import torch

from . import codec

DEFAULT_MODE = "repaired"


class DecodeFault(IndexError):
    """The decoder produced symbol -1 (the reference carries on with negative-index wraparound,
    cabac_compression.py:288-292); decoding stops at that symbol here."""


class ContextModel:
    """Parameter holder with the reference constructor (cabac_compression.py:66-76)."""

    def __init__(self, n_symbols=256, context_size=5, adaptation_rate=0.05):
        self.n_symbols = n_symbols
        self.context_size = context_size
        self.adaptation_rate = adaptation_rate
        self.context_models = {}
        self.context_counts = {}

    def is_fresh(self):
        return len(self.context_models) == 0 and len(self.context_counts) == 0


def _require_fresh(context_model):
    models = getattr(context_model, "context_models", None)
    if models is not None and len(models) != 0:
        raise NotImplementedError("the CUDA coder starts every stream from a fresh ContextModel; "
                                  "pre-populated models are not supported and there is no CPU fallback")


def _device(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("image_compression_2_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def raise_for_status(status, fault_index, what):
    """Translate a per-stream status word into the exception the reference raises."""
    status = int(status)
    if status == codec.STATUS_OK:
        return
    msg = "%s: %s at symbol %d" % (what, codec.STATUS_NAMES.get(status, status), int(fault_index))
    if status == 1:
        raise ValueError("byte must be in range(0, 256) [%s]" % msg)      # :193/:197
    if status == 2:
        raise IndexError("index out of bounds for cumulative_probs [%s]" % msg)  # :291
    if status == 3:
        raise ZeroDivisionError("float division by zero [%s]" % msg)      # :285
    if status == 4:
        raise DecodeFault(msg)
    if status == 6:
        raise IndexError("symbol index out of range [%s]" % msg)           # probs[symbol], :350
    raise RuntimeError(msg)


def cabac_encode(data, context_model, mode=None, device=None):
    """cabac_compression.py:315.  data: int array; a 3-D array (B,R,C) is ONE stream whose B images
    share the coder and model, any other rank uses a single global context.  Returns `bytes` holding
    one byte (0/1) per emitted bit, exactly what the reference returns (defect D1)."""
    packed, nbits = cabac_encode_packed(data, context_model, mode=mode, device=device)
    return np.unpackbits(np.frombuffer(packed, dtype=np.uint8))[:nbits].tobytes()


def cabac_encode_packed(data, context_model, mode=None, device=None):
    """Same stream as cabac_encode, returned as (MSB-first packed bytes, nbits)."""
    _require_fresh(context_model)
    dev = _device(device)
    arr = np.ascontiguousarray(np.asarray(data), dtype=np.int32)
    layout = codec.layout_reference(arr.shape)
    if layout.total == 0:
        raise ValueError("empty input")
    idx = torch.from_numpy(arr.reshape(-1)).to(dev)
    enc = codec.encode_batch(idx, layout, context_model.n_symbols, mode=mode or DEFAULT_MODE,
                             adaptation_rate=context_model.adaptation_rate)
    streams, nbits, status, fault = enc.to_host()
    if int(status[0]) == 5:  # slot too small for this stream: retry with the worst case
        enc = codec.encode_batch(idx, layout, context_model.n_symbols, mode=mode or DEFAULT_MODE,
                                 adaptation_rate=context_model.adaptation_rate, slot_bytes=layout.total * 12 + 256)
        streams, nbits, status, fault = enc.to_host()
    raise_for_status(status[0], fault[0], "cabac_encode")
    return streams[0], int(nbits[0])


def cabac_decode(encoded_bytes, context_model, shape, mode=None, device=None):
    """cabac_compression.py:363.  encoded_bytes: MSB-first PACKED bits (what the reference decoder
    reads, :260-270).  Returns int32 ndarray of `shape`."""
    _require_fresh(context_model)
    dev = _device(device)
    shape = tuple(int(s) for s in shape)
    layout = codec.layout_reference(shape)
    data, offsets, nbits = codec.pack_streams_for_device([bytes(encoded_bytes)], dev)
    idx, _, status, fault = codec.decode_batch(data, offsets, nbits, layout, context_model.n_symbols,
                                               mode=mode or DEFAULT_MODE,
                                               adaptation_rate=context_model.adaptation_rate)
    st, fi = int(status.cpu()[0]), int(fault.cpu()[0])
    raise_for_status(st, fi, "cabac_decode")
    return idx.cpu().numpy().reshape(shape)


# ---- batch extensions (not in the reference): B independent streams, fresh model each ----------

def cabac_encode_batch(data, n_symbols=256, adaptation_rate=0.05, mode=None, device=None):
    """data [B,R,C] int -> (list of packed bytes, nbits int32[B], status int32[B], fault int32[B])."""
    dev = _device(device)
    if isinstance(data, torch.Tensor):
        idx = data.to(dev, dtype=torch.int32).contiguous()
        shape = tuple(data.shape)
    else:
        arr = np.ascontiguousarray(np.asarray(data), dtype=np.int32)
        idx, shape = torch.from_numpy(arr).to(dev), arr.shape
    enc = codec.encode_batch(idx, codec.layout_independent(shape), n_symbols, mode=mode or DEFAULT_MODE,
                             adaptation_rate=adaptation_rate)
    return enc.to_host()


def cabac_decode_batch(streams, shape, n_symbols=256, adaptation_rate=0.05, mode=None, device=None):
    """streams: list of packed bytes (one per stream of shape[1:]) -> (int32 ndarray shape, status, fault)."""
    dev = _device(device)
    layout = codec.layout_independent(shape)
    data, offsets, nbits = codec.pack_streams_for_device(list(streams), dev)
    idx, _, status, fault = codec.decode_batch(data, offsets, nbits, layout, n_symbols, mode=mode or DEFAULT_MODE,
                                               adaptation_rate=adaptation_rate)
    return idx.cpu().numpy().reshape(shape), status.cpu().numpy(), fault.cpu().numpy()
