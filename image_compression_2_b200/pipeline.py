"""The whole hot path as one object: W+ latents -> quantise -> encode -> (bytes) -> decode -> dequantise.

This is what bench.py times and what a caller batching many images would use.  Entry points:
  roundtrip_device  inputs already in HBM, results left in HBM (the kernel-level metric)
  roundtrip_host    HOST buffers in and out: pinned fp32 latents are copied up, the compressed
                    streams are copied down (what save_compressed produces), copied up again for
                    decoding, and the dequantised fp32 latents are copied down (what
                    load_compressed hands to the generator).  The end-to-end metric, one batch at a time.
  roundtrip_host_stream  the same trip for a sequence of host batches with two of them in flight (the
                    copies of one batch under the kernels of the other): the end-to-end THROUGHPUT metric.
"""
import torch

from . import codec


class LatentPipeline:
    def __init__(self, n_symbols=256, R=16, C=512, quantizer="codebook", mode="repaired", adaptation_rate=0.05,
                 device=None, idx_dtype=None):
        if not torch.cuda.is_available():
            raise RuntimeError("LatentPipeline needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n = int(n_symbols)
        self.bits = self.n.bit_length() - 1
        self.R, self.C = int(R), int(C)
        self.quantizer = quantizer
        self.mode = mode
        self.rate = float(adaptation_rate)
        # indices travel between the kernels as uint8 (up to 256 symbols) or int16: 1-2 bytes of HBM traffic per
        # symbol instead of the int32 the reference's host code uses (idx_dtype=torch.int32 restores that)
        self.idx_dtype = codec.idx_dtype_for(self.n) if idx_dtype is None else idx_dtype
        # table from the host's torch.linspace, as the reference builds it
        # (gumbel_softmax_compression.py:49-52); for quantiser A the dequantisation table is the
        # affine grid idx/(2^bits-1)*2-1 produced by the dequantise-A kernel itself
        self.codebook = torch.linspace(-1, 1, self.n).float().to(self.device)
        if quantizer == "affine":
            grid = torch.arange(self.n, dtype=torch.int32, device=self.device)
            self.deq_table = codec.dequantize_affine(grid, self.bits)
        elif quantizer == "codebook":
            self.deq_table = self.codebook
        else:
            raise ValueError(quantizer)
        self.ws = codec.CoderWorkspace()
        self._pinned = {}
        self._bytes_hint = 0  # roundtrip_host: largest compressed stream seen so far, in bytes
        self._chunk_ctx = []  # roundtrip_host: (stream, workspace) per chunk
        self._stream_ctx = []  # roundtrip_host_stream: (stream, workspace) per slot

    # ---- stages -------------------------------------------------------------------------------
    def quantize(self, latents):
        if self.quantizer == "codebook":
            idx, _ = codec.quantize_codebook(latents, self.codebook, sorted_ascending=True, idx_dtype=self.idx_dtype)
        else:
            idx, _ = codec.quantize_affine(latents, self.bits, want_wq=False, idx_dtype=self.idx_dtype)
            # quantiser A does not clamp (stylegan3_hvae_full.py:313-316); the coder alphabet does (the narrow
            # index types are clamped by the kernel)
            if self.idx_dtype == torch.int32:
                idx = idx.clamp_(0, self.n - 1)
        return idx

    def encode(self, idx, ws=None, reuse_output=False):
        B = idx.shape[0]
        layout = codec.StreamLayout(B, 1, self.R, self.C, 1)
        return codec.encode_batch(idx.reshape(-1), layout, self.n, mode=self.mode, adaptation_rate=self.rate,
                                  workspace=ws or self.ws, reuse_output=reuse_output)

    def decode(self, data, offsets, nbits, B, ws=None, deq_out=None, want_idx=True, reuse_output=False, flags=0):
        layout = codec.StreamLayout(B, 1, self.R, self.C, 1)
        return codec.decode_batch(data, offsets, nbits, layout, self.n, mode=self.mode, adaptation_rate=self.rate,
                                  codebook=self.deq_table, workspace=ws or self.ws, deq_out=deq_out, flags=flags,
                                  idx_dtype=self.idx_dtype, want_idx=want_idx, reuse_output=reuse_output)

    # ---- whole path ---------------------------------------------------------------------------
    def roundtrip_device(self, latents):
        """latents fp32 [B,R,C] on the GPU. Returns dict with idx, enc (EncodedBatch), dec_idx, deq, statuses."""
        idx = self.quantize(latents)
        enc = self.encode(idx)
        dec_idx, deq, dstatus, dfault = self.decode(enc.data, enc.offsets, enc.nbits, latents.shape[0])
        return dict(idx=idx, enc=enc, dec_idx=dec_idx.view(latents.shape), deq=deq.view(latents.shape),
                    dec_status=dstatus, dec_fault=dfault)

    def _pin(self, key, shape, dtype):
        buf = self._pinned.get(key)
        if buf is None or buf.shape != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(tuple(shape), dtype=dtype).pin_memory()
            self._pinned[key] = buf
        return buf

    def roundtrip_host(self, latents_host, chunks=None):
        """latents_host: pinned CPU fp32 [B,R,C].  Returns dict(bytes = pinned uint8 buffer holding the compressed
        streams, offsets, nbits, enc_status, deq = pinned fp32 [B,R,C], dec_status, h2d_bytes, d2h_bytes).
        Synchronises once, at the end.  The returned tensors are pinned buffers of this object: the next call
        overwrites them (copy what has to outlive it).

        Every chunk makes the full trip -- latents up, compressed bytes down to the host (what save_compressed
        writes) and up again (what load_compressed reads), dequantised fp32 down (written by the decoder straight
        into the pinned buffer) -- and nothing in it waits for the host: the number of compressed bytes to move is
        not read back mid-way but bounded by a size hint (the largest per-stream size seen by earlier calls, +2 %;
        the whole slot area on the first call), and checked after the final synchronise; a chunk whose streams
        outgrew the hint is simply done again with the exact size.

        The batch is cut into `chunks` contiguous sub-batches (default: chunks of about 1400-2000 streams from 1536
        streams up), each with its own CUDA stream and workspace, all enqueued up front, so the copies of one chunk
        run under the kernels of the others.  Measured on a B200 (tools/e2e_probe.py): 4096 streams 33.2 ms in one
        piece, 31.2 / 30.2 / 32.9 ms in 2 / 3 / 4 chunks; 1024 streams 9.79 ms in one piece, 10.07 / 10.40 / 10.57 ms
        in 2 / 3 / 4 -- there every kernel is a single latency-bound wave and splitting it only adds waves."""
        B = latents_host.shape[0]
        if chunks is None:
            chunks = max(2, min(4, (B + 700) // 1400)) if B >= 1536 else 1
        chunks = max(1, min(int(chunks), B))
        bounds = [(B * c) // chunks for c in range(chunks + 1)]
        main = torch.cuda.current_stream(self.device)
        while len(self._chunk_ctx) < chunks:
            self._chunk_ctx.append((torch.cuda.Stream(device=self.device), codec.CoderWorkspace()))
        slot = int(codec._native.load().lc_encode_slot_bytes(1, self.R, self.C, self.n))
        per_stream = min(slot, int(self._bytes_hint * 1.02) + 64) if self._bytes_hint else slot
        deq_host = self._pin("deq", latents_host.shape, torch.float32)
        st_host = self._pin("dstatus", (B,), torch.int32)
        meta_all = self._pin("meta", (3 * B + chunks,), torch.int64)  # per chunk: offsets[Bc+1] | nbits[Bc] | status[Bc]
        caps = [((bounds[c + 1] - bounds[c]) * per_stream + 15) // 16 * 16 for c in range(chunks)]
        bytes_host = self._pin("bytes", (sum(caps),), torch.uint8)
        with torch.cuda.device(self.device):
            redo = self._roundtrip_chunks(latents_host, bounds, caps, range(chunks), main, deq_host, st_host, meta_all,
                                          bytes_host)
            if redo:  # streams larger than the hint: those chunks again, with room for any stream
                full = [((bounds[c + 1] - bounds[c]) * slot + 15) // 16 * 16 for c in range(chunks)]
                bytes_host = self._pin("bytes_full", (sum(full),), torch.uint8)
                self._roundtrip_chunks(latents_host, bounds, full, range(chunks), main, deq_host, st_host, meta_all,
                                       bytes_host)
                caps = full
        offs_out = torch.empty(B + 1, dtype=torch.int64)
        nbits_out = torch.empty(B, dtype=torch.int64)
        status_out = torch.empty(B, dtype=torch.int64)
        base, used_total = 0, 0
        for c in range(chunks):
            c0, c1 = bounds[c], bounds[c + 1]
            Bc = c1 - c0
            m = meta_all[3 * c0 + c:3 * c0 + c + 3 * Bc + 1]
            offs_out[c0:c1] = m[:Bc] + base
            nbits_out[c0:c1] = m[Bc + 1:2 * Bc + 1]
            status_out[c0:c1] = m[2 * Bc + 1:]
            used_total += int(m[Bc])
            base += caps[c]
        offs_out[B] = base
        ok = status_out == 0
        if bool(ok.any()):
            self._bytes_hint = max(self._bytes_hint, int(((nbits_out[ok] + 7) // 8).max()) + 16)
        moved = sum(caps)
        h2d = latents_host.numel() * 4 + moved + (2 * B + chunks) * 8
        d2h = meta_all.numel() * 8 + moved + deq_host.numel() * 4 + B * 4
        return dict(bytes=bytes_host, offsets=offs_out, nbits=nbits_out, enc_status=status_out, deq=deq_host,
                    dec_status=st_host, h2d_bytes=h2d, d2h_bytes=d2h, chunks=chunks, compressed_bytes=used_total)

    def _enqueue_trip(self, lat_host, st, ws, cap, bytes_seg, meta_host, deq_host, st_host, dec_flags=0):
        """Enqueue one batch's whole trip on CUDA stream `st` (no host wait): latents up, quantise, encode, compressed
        bytes down into `bytes_seg` (at most `cap`) and up again, decode with the dequantised rows written straight
        into the pinned `deq_host`.  meta_host receives offsets[Bc+1] | nbits[Bc] | status[Bc]."""
        Bc = lat_host.shape[0]
        with torch.cuda.stream(st):
            lat = lat_host.to(self.device, non_blocking=True)
            enc = self.encode(self.quantize(lat), ws, reuse_output=True)
            meta_dev = ws.get(("meta", self.device), 3 * Bc + 1, self.device, torch.int64)[:3 * Bc + 1]
            meta_dev[:Bc + 1] = enc.offsets
            meta_dev[Bc + 1:2 * Bc + 1] = enc.nbits
            meta_dev[2 * Bc + 1:] = enc.status
            meta_host.copy_(meta_dev, non_blocking=True)
            cap = min(cap, enc.data.numel())
            seg = bytes_seg[:cap]
            # compressed product -> host (what save_compressed would write) ...
            seg.copy_(enc.data[:cap], non_blocking=True)
            # ... and host -> device again (what load_compressed would read), decode + dequantise
            data_dev = ws.get(("bytes_back", self.device), cap, self.device)[:cap]
            data_dev.copy_(seg, non_blocking=True)
            meta_back = ws.get(("meta_back", self.device), 3 * Bc + 1, self.device, torch.int64)[:3 * Bc + 1]
            meta_back.copy_(meta_host, non_blocking=True)
            nbits_dev = ws.get(("nbits_back", self.device), Bc, self.device, torch.int32)[:Bc]
            nbits_dev.copy_(meta_back[Bc + 1:2 * Bc + 1])
            # (the decoder writes the dequantised rows straight into the pinned host buffer as it goes; a
            # stream cut short by the cap decodes garbage inside its own slot and is redone by the caller)
            _, _, dstatus, _ = self.decode(data_dev, meta_back[:Bc + 1], nbits_dev, Bc, ws, deq_out=deq_host,
                                           want_idx=False, reuse_output=True, flags=dec_flags)
            st_host.copy_(dstatus, non_blocking=True)

    def _roundtrip_chunks(self, latents_host, bounds, caps, which, main, deq_host, st_host, meta_all, bytes_host):
        """Enqueue the whole trip of the listed chunks (no host wait), synchronise, return the chunks whose compressed
        size exceeded their cap."""
        base = [sum(caps[:c]) for c in range(len(caps))]
        for c in which:
            c0, c1 = bounds[c], bounds[c + 1]
            Bc = c1 - c0
            st, ws = self._chunk_ctx[c]
            st.wait_stream(main)
            self._enqueue_trip(latents_host[c0:c1], st, ws, caps[c], bytes_host[base[c]:base[c] + caps[c]],
                               meta_all[3 * c0 + c:3 * c0 + c + 3 * Bc + 1], deq_host[c0:c1], st_host[c0:c1])
        for c in which:
            main.wait_stream(self._chunk_ctx[c][0])
        main.synchronize()
        redo = []
        for c in which:
            c0, c1 = bounds[c], bounds[c + 1]
            if int(meta_all[3 * c0 + c + (c1 - c0)]) > caps[c]:
                redo.append(c)
        return redo

    # ---- streaming: several batches in flight ---------------------------------------------------
    def roundtrip_host_stream(self, batches, depth=2, dec_flags=0):
        """Generator over results (same dict as roundtrip_host, in order) for an iterable of pinned host batches,
        with up to `depth` batches in flight: every batch makes the same full trip as in roundtrip_host, on its own
        CUDA stream with its own workspace and pinned output buffers, so the host->device copy of the next batch's
        latents (0.65 ms of the 9.5 ms a batch of 1024 takes on its own) runs under the kernels of the one before.
        Nothing waits for the host between batches; the generator synchronises only on the event of the batch it is
        about to hand out.  A result's buffers belong to its slot: they are overwritten `depth` batches later."""
        import collections
        pending = collections.deque()
        for k, lat in enumerate(batches):
            if len(pending) >= depth:
                yield self._stream_collect(pending.popleft())
            pending.append(self._stream_submit(lat, k % depth, dec_flags))
        while pending:
            yield self._stream_collect(pending.popleft())

    def _stream_submit(self, latents_host, slot, dec_flags=0):
        B = latents_host.shape[0]
        while len(self._stream_ctx) <= slot:
            self._stream_ctx.append((torch.cuda.Stream(device=self.device), codec.CoderWorkspace()))
        st, ws = self._stream_ctx[slot]
        slot_bytes = int(codec._native.load().lc_encode_slot_bytes(1, self.R, self.C, self.n))
        per_stream = min(slot_bytes, int(self._bytes_hint * 1.02) + 64) if self._bytes_hint else slot_bytes
        cap = (B * per_stream + 15) // 16 * 16
        bufs = dict(deq=self._pin(("s_deq", slot), latents_host.shape, torch.float32),
                    st=self._pin(("s_dstatus", slot), (B,), torch.int32),
                    meta=self._pin(("s_meta", slot), (3 * B + 1,), torch.int64),
                    bytes=self._pin(("s_bytes", slot, cap), (cap,), torch.uint8))
        with torch.cuda.device(self.device):
            st.wait_stream(torch.cuda.current_stream(self.device))
            self._enqueue_trip(latents_host, st, ws, cap, bufs["bytes"], bufs["meta"], bufs["deq"], bufs["st"], dec_flags)
            ev = torch.cuda.Event()
            ev.record(st)
        return dict(lat=latents_host, B=B, cap=cap, ev=ev, **bufs)

    def _stream_collect(self, tk):
        tk["ev"].synchronize()
        B, m = tk["B"], tk["meta"]
        if int(m[B]) > tk["cap"]:  # streams larger than the hint: this batch again, on its own, with room for any stream
            return self.roundtrip_host(tk["lat"], chunks=1)
        offs = m[:B + 1].clone()
        nbits, status = m[B + 1:2 * B + 1].clone(), m[2 * B + 1:].clone()
        ok = status == 0
        if bool(ok.any()):
            self._bytes_hint = max(self._bytes_hint, int(((nbits[ok] + 7) // 8).max()) + 16)
        lat = tk["lat"]
        return dict(bytes=tk["bytes"], offsets=offs, nbits=nbits, enc_status=status, deq=tk["deq"], dec_status=tk["st"],
                    h2d_bytes=lat.numel() * 4 + tk["cap"] + (2 * B + 1) * 8,
                    d2h_bytes=m.numel() * 8 + tk["cap"] + lat.numel() * 4 + B * 4, chunks=1,
                    compressed_bytes=int(m[B]))
