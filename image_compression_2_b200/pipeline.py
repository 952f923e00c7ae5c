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
                 device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("LatentPipeline needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n = int(n_symbols)
        self.bits = self.n.bit_length() - 1
        self.R, self.C = int(R), int(C)
        self.quantizer = quantizer
        self.mode = mode
        self.rate = float(adaptation_rate)
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

    # ---- stages -------------------------------------------------------------------------------
    def quantize(self, latents):
        if self.quantizer == "codebook":
            idx, _ = codec.quantize_codebook(latents, self.codebook, sorted_ascending=True)
        else:
            idx, _ = codec.quantize_affine(latents, self.bits, want_wq=False)
            # quantiser A does not clamp (stylegan3_hvae_full.py:313-316); the coder alphabet does
            idx = idx.clamp_(0, self.n - 1)
        return idx

    def encode(self, idx):
        B = idx.shape[0]
        layout = codec.StreamLayout(B, 1, self.R, self.C, 1)
        return codec.encode_batch(idx.reshape(-1), layout, self.n, mode=self.mode, adaptation_rate=self.rate,
                                  workspace=self.ws)

    def decode(self, data, offsets, nbits, B):
        layout = codec.StreamLayout(B, 1, self.R, self.C, 1)
        return codec.decode_batch(data, offsets, nbits, layout, self.n, mode=self.mode, adaptation_rate=self.rate,
                                  codebook=self.deq_table, workspace=self.ws)

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

    def roundtrip_host(self, latents_host):
        """latents_host: pinned CPU fp32 [B,R,C].  Returns (streams_host uint8 pinned view, offsets, nbits,
        deq_host fp32 pinned [B,R,C], h2d_bytes, d2h_bytes).  Synchronises at the end."""
        B = latents_host.shape[0]
        lat = latents_host.to(self.device, non_blocking=True)
        idx = self.quantize(lat)
        enc = self.encode(idx)
        # compressed product -> host (what save_compressed would write)
        meta_dev = torch.cat([enc.offsets, enc.nbits.long(), enc.status.long()])
        meta_host = self._pin("meta", meta_dev.shape, torch.int64)
        meta_host.copy_(meta_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        used = int(meta_host[B])
        bytes_host = self._pin("bytes", (enc.data.numel(),), torch.uint8)
        bytes_host[:used].copy_(enc.data[:used], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        # host -> device again (what load_compressed would read), decode + dequantise
        data_dev = bytes_host[:max(used, 16)].to(self.device, non_blocking=True)
        offs_dev = meta_host[:B + 1].to(self.device, non_blocking=True)
        nbits_dev = meta_host[B + 1:2 * B + 1].to(self.device, non_blocking=True).int()
        dec_idx, deq, dstatus, dfault = self.decode(data_dev, offs_dev, nbits_dev, B)
        deq_host = self._pin("deq", latents_host.shape, torch.float32)
        deq_host.copy_(deq.view(latents_host.shape), non_blocking=True)
        st_host = self._pin("dstatus", (B,), torch.int32)
        st_host.copy_(dstatus, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h2d = latents_host.numel() * 4 + used + (2 * B + 1) * 8
        d2h = meta_host.numel() * 8 + used + deq_host.numel() * 4 + B * 4
        return dict(bytes=bytes_host[:used], offsets=meta_host[:B + 1], nbits=meta_host[B + 1:2 * B + 1],
                    enc_status=meta_host[2 * B + 1:], deq=deq_host, dec_status=st_host, h2d_bytes=h2d, d2h_bytes=d2h)
