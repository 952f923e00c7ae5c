"""Drop-in host surface for the reference's coder functions, backed by the CUDA kernels.

Mirrors /root/reference/cabac_compression.py: ContextModel (:60-162, constructor and state
containers only -- the model itself lives on the GPU), cabac_encode (:315-359), cabac_decode
(:363-406).  Same argument meaning, same return types, same exception classes.

Deviations, all documented in DESIGN.md:
  * a FRESH (empty) ContextModel takes the fast kernels and is NOT filled in after the call, unless it was
    created with track_state=True;
  * a non-empty model (or track_state=True) takes the stateful kernel (csrc/lc_stateful.cuh): the stream is coded
    from that model and the object is left holding what the reference's object would hold (the reference mutates
    one shared model across calls, defect D5).  Alphabets up to 1024 symbols (the direct-mapped device table is
    (n+1)^2*n*8 bytes: 8.6 GB of the B200's 180 GB at 1024 symbols; only valid[]/counts[] are cleared);
  * DEFAULT_MODE is "repaired": the file as shipped ("verbatim") raises ValueError within a few
    hundred symbols on any realistic stream (defect D3).  mode="verbatim" reproduces that.
"""
import numpy as np
import torch

from . import codec

DEFAULT_MODE = "repaired"


class DecodeFault(IndexError):
    """Status 4 of ABI version 1 (decoding stopped at symbol -1).  Not raised any more: the kernels now follow the
    reference through symbol -1 (NumPy negative indexing, cabac_compression.py:288-292,403) and the decoded array
    simply holds -1 there, as the reference's does."""


class ContextModel:
    """The reference's model object (cabac_compression.py:60-162): same constructor, same two state containers,
    same three methods.  The coder kernels keep their own device-side model while a stream is being coded; this
    object is the host view of it -- read by a stateful call (import), filled in after it (export).  The methods
    work on that view: `get_context` and `get_probability` are index arithmetic and a dictionary lookup,
    `update_model` runs ContextModel.update_model on the GPU (lc_model_update: the same exact float64 update,
    NumPy pairwise-sum order, the coder kernels apply) -- there is no CPU arithmetic path here either."""

    def __init__(self, n_symbols=256, context_size=5, adaptation_rate=0.05, track_state=False):
        self.n_symbols = n_symbols
        self.context_size = context_size
        self.adaptation_rate = adaptation_rate
        self.context_models = {}   # (left, up) or () -> float64[n_symbols], as the reference keeps them
        self.context_counts = {}   # ... -> number of update_model calls
        self.track_state = track_state  # extension: fill the two dicts in even when the model starts empty

    def is_fresh(self):
        return len(self.context_models) == 0 and len(self.context_counts) == 0

    def get_context(self, data, pos, shape):
        """(:78-117) (left, up) neighbours with -1 sentinels for a 3-D shape, () otherwise."""
        if len(shape) != 3:
            return ()
        batch, ws, dim = np.unravel_index(pos, shape)
        left = int(data[batch, ws, dim - 1]) if dim > 0 else -1
        up = int(data[batch, ws - 1, dim]) if ws > 0 else -1
        return (left, up)

    def get_probability(self, context, symbol=None):
        """(:146-162) the context's vector (a missing context is ones(n)/n and exists from now on, :73)."""
        context = tuple(int(c) for c in context)
        probs = self.context_models.get(context)
        if probs is None:
            probs = self.context_models[context] = np.ones(self.n_symbols) / self.n_symbols
        return probs if symbol is None else probs[symbol]

    def update_model(self, context, symbol, device=None):
        """(:119-144) EMA on `symbol`, the others rescaled by the pairwise-sum factor; counts += 1."""
        from . import _native
        lib = _native.load()
        dev = _device(device)
        context = tuple(int(c) for c in context)
        n = int(self.n_symbols)
        symbol = int(symbol)
        if not (-n <= symbol < n):
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (symbol, n))
        vec = torch.from_numpy(np.ascontiguousarray(self.get_probability(context), np.float64)).to(dev)
        with codec._On(vec) as on:
            _native.check(lib.lc_model_update(vec.data_ptr(), n, symbol, float(self.adaptation_rate), on.stream),
                          "lc_model_update")
        self.context_models[context] = vec.cpu().numpy()
        self.context_counts[context] = self.context_counts.get(context, 0) + 1


def _wants_state(context_model):
    models = getattr(context_model, "context_models", None)
    return bool(getattr(context_model, "track_state", False)) or (models is not None and len(models) != 0)


class _StatefulTable:
    """Device table of csrc/lc_stateful.cuh for one call: the given model scattered in, the result gathered out."""

    def __init__(self, context_model, layout, dev):
        lib = _native_lib()
        n, has_ctx = int(context_model.n_symbols), int(layout.has_ctx)
        nbytes = lib.lc_stateful_table_bytes(n, has_ctx)
        if nbytes < 0:
            raise NotImplementedError("stateful coding needs a power-of-two alphabet of at most 1024 symbols; "
                                      "there is no CPU fallback")
        self.n, self.has_ctx, self.nkeys = n, has_ctx, ((n + 1) ** 2 if has_ctx else 1)
        off_c, off_v = lib.lc_stateful_table_offset(n, has_ctx, 1), lib.lc_stateful_table_offset(n, has_ctx, 2)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.buf[:off_v].zero_()
        self.valid = self.buf[:self.nkeys]
        self.counts = self.buf[off_c:off_c + 4 * self.nkeys].view(torch.int32)
        self.vecs = self.buf[off_v:off_v + 8 * self.nkeys * n].view(torch.float64).view(self.nkeys, n)
        models = getattr(context_model, "context_models", {}) or {}
        if len(models):
            keys = list(models.keys())
            idx = torch.tensor([self._index(k) for k in keys], dtype=torch.long, device=dev)
            vecs = np.stack([np.asarray(models[k], np.float64).reshape(n) for k in keys])
            counts = getattr(context_model, "context_counts", {}) or {}
            self.vecs[idx] = torch.from_numpy(vecs).to(dev)
            self.valid[idx] = 1
            self.counts[idx] = torch.tensor([int(counts.get(k, 0)) for k in keys], dtype=torch.int32, device=dev)

    def _index(self, key):
        if not self.has_ctx:
            if len(key) != 0:
                raise ValueError("model keys must be () for non-3-D data, got %r" % (key,))
            return 0
        left, up = int(key[0]), int(key[1])
        if not (-1 <= left < self.n and -1 <= up < self.n):
            raise ValueError("context %r outside the alphabet" % (key,))
        return (left + 1) * (self.n + 1) + (up + 1)

    def export(self, context_model):
        """Leave in context_model what the reference object holds after the call (cabac_compression.py:143-144,157)."""
        idx = torch.nonzero(self.valid).reshape(-1)
        vecs = self.vecs[idx].cpu().numpy()
        counts = self.counts[idx].cpu().numpy()
        for j, k in enumerate(idx.cpu().numpy().tolist()):
            key = ((k // (self.n + 1)) - 1, (k % (self.n + 1)) - 1) if self.has_ctx else ()
            context_model.context_models[key] = vecs[j].copy()
            if counts[j] > 0:
                context_model.context_counts[key] = int(counts[j])


def _native_lib():
    from . import _native
    return _native.load()


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
    dev = _device(device)
    arr = np.ascontiguousarray(np.asarray(data), dtype=np.int32)
    layout = codec.layout_reference(arr.shape)
    if layout.total == 0:
        raise ValueError("empty input")
    idx = torch.from_numpy(arr.reshape(-1)).to(dev)
    if _wants_state(context_model):
        from . import _native
        lib = _native.load()
        table = _StatefulTable(context_model, layout, dev)
        slot_bytes = (layout.total * 12 + 256 + 15) // 16 * 16
        slot = torch.empty(slot_bytes, dtype=torch.uint8, device=dev)
        res = torch.zeros(3, dtype=torch.int32, device=dev)
        with codec._On(idx, table.buf, slot, res) as on:
            _native.check(lib.lc_stateful_encode(idx.data_ptr(), layout.imgs, layout.R, layout.C,
                                                 int(context_model.n_symbols), float(context_model.adaptation_rate),
                                                 codec.MODES[mode or DEFAULT_MODE], layout.has_ctx, table.buf.data_ptr(),
                                                 table.buf.numel(), slot.data_ptr(), slot_bytes, res[0:].data_ptr(),
                                                 res[1:].data_ptr(), res[2:].data_ptr(), on.stream), "lc_stateful_encode")
        nb, st, fi = (int(x) for x in res.cpu())
        table.export(context_model)  # the reference has mutated the model up to the fault before it raises
        raise_for_status(st, fi, "cabac_encode")
        return slot[:(nb + 7) // 8].cpu().numpy().tobytes(), nb
    # (a stream that overflows the default slot is coded again with the worst-case slot)
    _, (streams, nbits, status, fault) = codec.encode_batch_checked(idx, layout, context_model.n_symbols,
                                                                    mode=mode or DEFAULT_MODE,
                                                                    adaptation_rate=context_model.adaptation_rate)
    raise_for_status(status[0], fault[0], "cabac_encode")
    return streams[0], int(nbits[0])


def cabac_decode(encoded_bytes, context_model, shape, mode=None, device=None):
    """cabac_compression.py:363.  encoded_bytes: MSB-first PACKED bits (what the reference decoder
    reads, :260-270).  Returns int32 ndarray of `shape`."""
    dev = _device(device)
    shape = tuple(int(s) for s in shape)
    layout = codec.layout_reference(shape)
    data, offsets, nbits = codec.pack_streams_for_device([bytes(encoded_bytes)], dev)
    if _wants_state(context_model):
        from . import _native
        lib = _native.load()
        table = _StatefulTable(context_model, layout, dev)
        out = torch.empty(layout.total, dtype=torch.int32, device=dev)
        res = torch.zeros(2, dtype=torch.int32, device=dev)
        with codec._On(data, table.buf, out, res) as on:
            _native.check(lib.lc_stateful_decode(data.data_ptr(), len(bytes(encoded_bytes)), layout.imgs, layout.R,
                                                 layout.C, int(context_model.n_symbols),
                                                 float(context_model.adaptation_rate), codec.MODES[mode or DEFAULT_MODE],
                                                 layout.has_ctx, table.buf.data_ptr(), table.buf.numel(), out.data_ptr(),
                                                 res[0:].data_ptr(), res[1:].data_ptr(), on.stream), "lc_stateful_decode")
        st, fi = (int(x) for x in res.cpu())
        table.export(context_model)
        raise_for_status(st, fi, "cabac_decode")
        return out.cpu().numpy().reshape(shape)
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
    _, host = codec.encode_batch_checked(idx, codec.layout_independent(shape), n_symbols, mode=mode or DEFAULT_MODE,
                                         adaptation_rate=adaptation_rate)
    return host


def cabac_decode_batch(streams, shape, n_symbols=256, adaptation_rate=0.05, mode=None, device=None):
    """streams: list of packed bytes (one per stream of shape[1:]) -> (int32 ndarray shape, status, fault)."""
    dev = _device(device)
    layout = codec.layout_independent(shape)
    data, offsets, nbits = codec.pack_streams_for_device(list(streams), dev)
    idx, _, status, fault = codec.decode_batch(data, offsets, nbits, layout, n_symbols, mode=mode or DEFAULT_MODE,
                                               adaptation_rate=adaptation_rate)
    return idx.cpu().numpy().reshape(shape), status.cpu().numpy(), fault.cpu().numpy()
