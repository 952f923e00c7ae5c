"""Drop-in compressor classes: the reference's Python call surface over the CUDA hot path.

Signatures follow the reference:
  StyleGAN3Compressor       /root/reference/stylegan3_hvae_full.py:250-380
  GumbelSoftmaxDiscretization /root/reference/gumbel_softmax_compression.py:26-137 (hot-path part)
  GumbelSoftmaxCompressor   /root/reference/gumbel_softmax_compression.py:140-319
  CABACCompressor           /root/reference/cabac_compression.py:409-588

The encoder (HVAE_VGG_Encoder) and the StyleGAN3 generator are passed in by the caller and stay on
the reference's PyTorch code -- they are out of scope.  Everything between `means` and
`generator.synthesis(w)` runs in liblatentcodec.so on the GPU the latents live on; nothing here
computes on the CPU, and a CPU tensor is refused rather than silently processed.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import codec, coder, containers


def _means_of(encoder, x, deterministic=True):
    w_plus, means, _ = encoder(x)
    return (means if deterministic else w_plus), w_plus


def _match_size(img, x, training_resolution):
    if training_resolution is not None and img.shape[2] != x.shape[2]:
        img = F.interpolate(img, size=(x.shape[2], x.shape[3]), mode="bilinear", align_corners=False)
    return img


def _module_device(module, default=None):
    for p in module.parameters():
        return p.device
    for b in module.buffers():
        return b.device
    return default


class StyleGAN3Compressor(nn.Module):
    """Quantiser A path: affine `quantization_bits` rounding of W+ and the .npz container."""

    def __init__(self, encoder, generator, training_resolution=None):
        super().__init__()
        self.encoder = encoder
        self.generator = generator
        self.training_resolution = training_resolution
        for p in generator.parameters():
            p.requires_grad = False

    def forward(self, x, noise_mode="const"):
        w_plus, _, _ = self.encoder(x)
        img = self.generator.synthesis(w_plus, noise_mode=noise_mode)
        return _match_size(img, x, self.training_resolution), w_plus

    def encode(self, x, deterministic=False):
        return _means_of(self.encoder, x, deterministic)[0]

    def compress(self, x, quantization_bits=8, deterministic=True):
        """-> dequantised fp32 W+ [B,num_ws,w_dim] (stylegan3_hvae_full.py:295-318), one fused kernel."""
        w, _ = _means_of(self.encoder, x, deterministic)
        _, wq = codec.quantize_affine(w.detach().float().contiguous(), quantization_bits, want_idx=False)
        return wq

    def compress_indices(self, x, quantization_bits=8, deterministic=True):
        """Extension: the integer indices the reference never materialises, plus the dequantised values."""
        w, _ = _means_of(self.encoder, x, deterministic)
        return codec.quantize_affine(w.detach().float().contiguous(), quantization_bits)

    def decompress(self, w_plus, noise_mode="const"):
        return self.generator.synthesis(w_plus, noise_mode=noise_mode)

    def save_compressed(self, x, filename, quantization_bits=8, deterministic=True):
        """stylegan3_hvae_full.py:331-361: same six .npz members, same return tuple."""
        w_q = self.compress(x, quantization_bits, deterministic)
        orig_size = x.numel() * 4
        comp_size = w_q.numel() * (quantization_bits / 8)
        containers.write_latent_npz(filename, w_q.detach().cpu().numpy(), x.shape[2:4], quantization_bits, orig_size,
                                    comp_size)
        return orig_size, comp_size, orig_size / comp_size

    def load_compressed(self, filename, noise_mode="const"):
        data = np.load(filename)
        w_q = torch.tensor(data["w"]).to(_module_device(self.generator))
        with torch.no_grad():
            img = self.decompress(w_q, noise_mode=noise_mode)
        return img, data["compression_ratio"]


class GumbelSoftmaxDiscretization(nn.Module):
    """Codebook quantiser (quantiser B), gumbel_softmax_compression.py:26-137.

    Same constructor, parameters and buffers as the reference layer (`codebook`, `log_temperature` -- a Parameter
    when learnable_temp --, `usage`), so a reference `discretization_state_dict` loads (cabac_compression.py:642-675),
    and the same accessors (`temperature`, `update_temp`, `get_code_usage`).  forward() provides the inference
    result the compressors consume: nearest-codebook indices (argmin, :118) and their codebook values, computed by
    the CUDA kernel without the [N,n] distance matrix.  The random Gumbel-softmax sample of the reference's
    training forward (:103-108) is training code and out of scope: `discretized` here is always codebook[argmin],
    the deterministic value `compress`/`decompress` use."""

    def __init__(self, latent_dim=512, n_embeddings=256, temperature=1.0, straight_through=True, learnable_temp=True):
        super().__init__()
        self.latent_dim = latent_dim
        self.n_embeddings = n_embeddings
        self.initial_temp = temperature
        self.straight_through = straight_through
        # the table comes from the host's torch.linspace, exactly as in the reference (:49-52);
        # the kernels take it as an argument and never rebuild it
        self.register_buffer("codebook", torch.linspace(-1, 1, n_embeddings).float())
        if learnable_temp:  # (:54-58)
            self.log_temperature = nn.Parameter(torch.ones(1) * np.log(temperature))
        else:
            self.register_buffer("log_temperature", torch.ones(1) * np.log(temperature))
        self.register_buffer("usage", torch.zeros(n_embeddings))  # (:61)

    @property
    def temperature(self):
        return torch.exp(self.log_temperature)

    def update_temp(self, anneal_rate=0.00003, min_temp=0.5):
        """(:67-71)"""
        with torch.no_grad():
            self.log_temperature.clamp_(min=np.log(min_temp))
            self.log_temperature -= anneal_rate

    def forward(self, z, hard=None):
        """-> (codebook[idx] shaped like z, perplexity of the hard assignment, idx int64 flat)."""
        idx, deq = codec.quantize_codebook(z.detach().float().contiguous(), self.codebook, want_deq=True)
        flat = idx.reshape(-1).long()
        counts = torch.bincount(flat, minlength=self.n_embeddings).float()
        if self.training:  # usage statistics (:121-123)
            self.usage += counts[: self.n_embeddings].to(self.usage.device)
        probs = counts / counts.sum().clamp(min=1)
        perplexity = torch.exp(-torch.sum(probs * torch.log(probs + 1e-10)))
        return deq, perplexity, flat

    def get_code_usage(self):
        """(:131-137)"""
        total = self.usage.sum().float()
        return self.usage / total if total > 0 else self.usage


class GumbelSoftmaxCompressor(nn.Module):
    def __init__(self, encoder, generator, n_embeddings=256, temperature=1.0, straight_through=True,
                 training_resolution=None):
        super().__init__()
        self.encoder = encoder
        self.generator = generator
        self.training_resolution = training_resolution
        self.discretization = GumbelSoftmaxDiscretization(latent_dim=getattr(encoder, "w_dim", 512),
                                                          n_embeddings=n_embeddings, temperature=temperature,
                                                          straight_through=straight_through)
        for p in generator.parameters():
            p.requires_grad = False

    def forward(self, x, noise_mode="const"):
        w_plus, means, _ = self.encoder(x)
        w_discrete, perplexity, _ = self.discretization(means)
        img = self.generator.synthesis(w_discrete, noise_mode=noise_mode)
        return _match_size(img, x, self.training_resolution), w_plus, w_discrete, perplexity

    def encode(self, x, deterministic=True):
        w, _ = _means_of(self.encoder, x, deterministic)
        return self.discretization(w, hard=True)[0]

    def compress_device(self, x):
        """Extension: int32 codes [B,num_ws,w_dim] left on the GPU (no host round trip)."""
        with torch.no_grad():
            _, means, _ = self.encoder(x)
            idx, _ = codec.quantize_codebook(means.detach().float().contiguous(),
                                             self.discretization.codebook.to(means.device))
        return idx

    def compress(self, x, discrete_bits=8):
        """gumbel_softmax_compression.py:213-235: int64 codes [B,num_ws,w_dim] on the CPU."""
        return self.compress_device(x).long().cpu()

    def decompress(self, codes, noise_mode="const"):
        with torch.no_grad():
            dev = self.discretization.codebook.device
            idx = codes.to(dev, dtype=torch.int32).contiguous()
            w_discrete = codec.dequantize_codebook(idx, self.discretization.codebook)
            return self.generator.synthesis(w_discrete, noise_mode=noise_mode)

    def save_compressed(self, x, filename, discrete_bits=8):
        codes_np = self.compress(x, discrete_bits=discrete_bits).numpy()
        orig_size = x.numel() * 4
        comp_size = codes_np.size * (np.log2(self.discretization.n_embeddings) / 8)
        containers.write_codes_npz(filename, codes_np, self.discretization.n_embeddings, x.shape[2:4], orig_size,
                                   comp_size)
        return orig_size, comp_size, orig_size / comp_size

    def load_compressed(self, filename, noise_mode="const"):
        data = np.load(filename)
        img = self.decompress(torch.from_numpy(data["codes"]), noise_mode=noise_mode)
        return img, data["compression_ratio"]


class CABACCompressor:
    """Codebook quantiser + context-adaptive arithmetic coder (cabac_compression.py:409-588).

    container="packed" (default) makes save_compressed/load_compressed round-trip: MSB-first packed
    bits and a correct header.  container="reference" reproduces the reference's output byte for
    byte (one byte per bit, comp_size counted in bits, header = 6; defects D1/D4), which its own
    loader cannot read back.  Every call starts from a fresh context model (defect D5 not kept) unless
    shared_model=True: then, as in the reference (:438,478,517), ONE ContextModel is mutated by every compress and
    decompress call (the stateful kernel, csrc/lc_stateful.cuh) -- a decompress after a compress then starts from
    the encoder's final model and does not reproduce the latents, exactly as the reference behaves.
    mode: "repaired" (default) or "verbatim" -- see image_compression_2_b200.coder."""

    def __init__(self, encoder, generator, discretization=None, n_embeddings=256, training_resolution=None,
                 container="packed", mode=None, shared_model=False):
        self.encoder = encoder
        self.generator = generator
        self.training_resolution = training_resolution
        if discretization is None:
            discretization = GumbelSoftmaxDiscretization(latent_dim=getattr(encoder, "w_dim", 512),
                                                         n_embeddings=n_embeddings)
        self.discretization = discretization
        self.context_model = coder.ContextModel(n_symbols=n_embeddings, track_state=bool(shared_model))
        self.shared_model = bool(shared_model)
        self.container = container
        self.mode = mode

    def encode(self, x, deterministic=True):
        w, _ = _means_of(self.encoder, x, deterministic)
        idx, deq = codec.quantize_codebook(w.detach().float().contiguous(), self.discretization.codebook.to(w.device),
                                           want_deq=True)
        return deq

    def _codes_device(self, x):
        with torch.no_grad():
            _, means, _ = self.encoder(x)
            idx, _ = codec.quantize_codebook(means.detach().float().contiguous(),
                                             self.discretization.codebook.to(means.device))
        return idx

    def compress(self, x, use_cabac=True):
        """-> (encoded_bytes, metadata) with the reference's metadata keys (:486-493)."""
        idx = self._codes_device(x)
        n = self.discretization.n_embeddings
        shape = tuple(idx.shape)
        orig_size = idx.numel() * np.log2(n) / 8
        if use_cabac and self.shared_model:
            packed, nb = coder.cabac_encode_packed(idx.cpu().numpy(), self.context_model, mode=self.mode, device=idx.device)
            encoded = (np.unpackbits(np.frombuffer(packed, dtype=np.uint8))[:nb].tobytes()
                       if self.container == "reference" else packed)
        elif use_cabac:
            layout = codec.layout_reference(shape)
            _, (streams, nbits, status, fault) = codec.encode_batch_checked(
                idx.reshape(-1), layout, n, mode=self.mode or coder.DEFAULT_MODE,
                adaptation_rate=self.context_model.adaptation_rate)
            coder.raise_for_status(status[0], fault[0], "CABACCompressor.compress")
            if self.container == "reference":
                encoded = np.unpackbits(np.frombuffer(streams[0], dtype=np.uint8))[: int(nbits[0])].tobytes()
            else:
                encoded = streams[0]
        else:
            encoded = idx.cpu().numpy().astype(np.int32).tobytes()
        comp_size = len(encoded)
        metadata = {"shape": shape, "n_embeddings": n, "use_cabac": use_cabac, "orig_size": orig_size,
                    "comp_size": comp_size, "compression_ratio": orig_size / comp_size}
        return encoded, metadata

    def decompress_latents(self, encoded_bytes, metadata):
        """Extension: decode + fused codebook lookup, result [B,num_ws,w_dim] fp32 on the GPU."""
        shape = tuple(metadata["shape"])
        n = metadata.get("n_embeddings", self.discretization.n_embeddings)
        dev = _module_device(self.generator, None) or self.discretization.codebook.device
        cb = self.discretization.codebook.to(dev)
        if metadata.get("use_cabac", True) and self.shared_model:
            codes = coder.cabac_decode(bytes(encoded_bytes), self.context_model, shape, mode=self.mode, device=dev)
            return codec.dequantize_codebook(torch.from_numpy(codes).to(dev), cb)
        if metadata.get("use_cabac", True):
            layout = codec.layout_reference(shape)
            data, offsets, nbits = codec.pack_streams_for_device([bytes(encoded_bytes)], dev)
            idx, deq, status, fault = codec.decode_batch(data, offsets, nbits, layout, n,
                                                         mode=self.mode or coder.DEFAULT_MODE,
                                                         adaptation_rate=self.context_model.adaptation_rate,
                                                         codebook=cb)
            coder.raise_for_status(int(status.cpu()[0]), int(fault.cpu()[0]), "CABACCompressor.decompress")
            return deq.reshape(shape)
        codes = torch.from_numpy(np.frombuffer(encoded_bytes, dtype=np.int32).reshape(shape).copy()).to(dev)
        return codec.dequantize_codebook(codes, cb)

    def decompress(self, encoded_bytes, metadata, noise_mode="const"):
        with torch.no_grad():
            w_discrete = self.decompress_latents(encoded_bytes, metadata)
            return self.generator.synthesis(w_discrete, noise_mode=noise_mode)

    def save_compressed(self, x, filename, use_cabac=True):
        encoded, metadata = self.compress(x, use_cabac=use_cabac)
        containers.write_cabac(filename, encoded, metadata, flavour=self.container)
        return metadata["orig_size"], metadata["comp_size"], metadata["compression_ratio"]

    def load_compressed(self, filename, noise_mode="const"):
        encoded, metadata = containers.read_cabac(filename)
        img = self.decompress(encoded, metadata, noise_mode=noise_mode)
        return img, metadata["compression_ratio"]
