"""The whole hot path as one object: W+ latents -> quantise -> encode -> (bytes) -> decode -> dequantise.

This is what bench.py times and what a caller batching many images would use.  Two entry points:
  roundtrip_device  inputs already in HBM, results left in HBM (the kernel-level metric)
  roundtrip_host    HOST buffers in and out: pinned fp32 latents are copied up, the compressed
                    streams are copied down (what save_compressed produces), copied up again for
                    decoding, and the dequantised fp32 latents are copied down (what
                    load_compressed hands to the generator).  The end-to-end metric.
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
        self._chunk_ctx = []  # roundtrip_host: (stream, workspace) per chunk

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
        """latents_host: pinned CPU fp32 [B,R,C].  Returns dict(bytes = pinned uint8 view of the compressed streams,
        offsets, nbits, enc_status, deq = pinned fp32 [B,R,C], dec_status, h2d_bytes, d2h_bytes).  Synchronises at
        the end.

        The batch is cut into `chunks` contiguous sub-batches (default: about 2048 streams each, from 1536 streams up), each with its own CUDA
        stream and workspace, so that the host<->device copies of one chunk run under the kernels of the others.
        Every chunk still makes the full trip: latents up, compressed bytes down to the host and up again, dequantised
        fp32 down.  Measured on a B200: 8192 streams 73.0 -> 63.5 ms; at 1024 streams (one latency-bound wave per
        kernel, encoder blocks of one chunk waiting for registers held by the decoder blocks of another) chunking
        gains nothing, so small batches stay in one piece."""
        B = latents_host.shape[0]
        if chunks is None:  # chunks of about 2048 streams, from 1536 streams up (measured: 1536 -6 %, 2048 -8 %, 4096 -6 %)
            chunks = max(2, min(4, (B + 1024) // 2048)) if B >= 1536 else 1
        chunks = max(1, min(int(chunks), B))
        bounds = [(B * c) // chunks for c in range(chunks + 1)]
        main = torch.cuda.current_stream()
        while len(self._chunk_ctx) < chunks:
            self._chunk_ctx.append((torch.cuda.Stream(device=self.device), codec.CoderWorkspace()))
        deq_host = self._pin("deq", latents_host.shape, torch.float32)
        st_host = self._pin("dstatus", (B,), torch.int32)
        meta_all = self._pin("meta", (3 * B + chunks,), torch.int64)  # per chunk: offsets[Bc+1] | nbits[Bc] | status[Bc]
        stage = []
        for c in range(chunks):  # everything up to the compressed product, all chunks enqueued before any host wait
            c0, c1 = bounds[c], bounds[c + 1]
            st, ws = self._chunk_ctx[c]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                lat = latents_host[c0:c1].to(self.device, non_blocking=True)
                enc = self.encode(self.quantize(lat), ws)
                m0 = 3 * c0 + c
                meta_host = meta_all[m0:m0 + 3 * (c1 - c0) + 1]
                meta_host.copy_(torch.cat([enc.offsets, enc.nbits.long(), enc.status.long()]), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(st)
            stage.append((enc, meta_host, ev))
        bytes_host = self._pin("bytes", (sum(e.data.numel() + 16 for e, _, _ in stage),), torch.uint8)
        base, used_total = 0, 0
        offs_out = torch.empty(B + 1, dtype=torch.int64)
        for c in range(chunks):
            c0, c1 = bounds[c], bounds[c + 1]
            Bc = c1 - c0
            st, ws = self._chunk_ctx[c]
            enc, meta_host, ev = stage[c]
            ev.synchronize()  # this chunk's sizes are on the host; the other chunks keep the GPU busy meanwhile
            used = int(meta_host[Bc])
            offs_out[c0:c1] = meta_host[:Bc] + base
            with torch.cuda.stream(st):
                # compressed product -> host (what save_compressed would write) ...
                bytes_host[base:base + used].copy_(enc.data[:used], non_blocking=True)
                # ... and host -> device again (what load_compressed would read), decode + dequantise
                data_dev = bytes_host[base:base + max(used, 16)].to(self.device, non_blocking=True)
                offs_dev = meta_host[:Bc + 1].to(self.device, non_blocking=True)
                nbits_dev = meta_host[Bc + 1:2 * Bc + 1].to(self.device, non_blocking=True).int()
                # (the decoder writes the dequantised rows straight into the pinned host buffer as it goes)
                dec_idx, deq, dstatus, dfault = self.decode(data_dev, offs_dev, nbits_dev, Bc, ws, deq_out=deq_host[c0:c1])
                st_host[c0:c1].copy_(dstatus, non_blocking=True)
            base += (used + 15) // 16 * 16
            used_total += used
        offs_out[B] = base
        for c in range(chunks):
            main.wait_stream(self._chunk_ctx[c][0])
        main.synchronize()
        nbits_out = torch.cat([m[(bounds[c + 1] - bounds[c]) + 1:2 * (bounds[c + 1] - bounds[c]) + 1] for c, (_, m, _) in enumerate(stage)])
        status_out = torch.cat([m[2 * (bounds[c + 1] - bounds[c]) + 1:] for c, (_, m, _) in enumerate(stage)])
        h2d = latents_host.numel() * 4 + used_total + (2 * B + chunks) * 8
        d2h = meta_all.numel() * 8 + used_total + deq_host.numel() * 4 + B * 4
        return dict(bytes=bytes_host[:base], offsets=offs_out, nbits=nbits_out, enc_status=status_out, deq=deq_host,
                    dec_status=st_host, h2d_bytes=h2d, d2h_bytes=d2h, chunks=chunks)
