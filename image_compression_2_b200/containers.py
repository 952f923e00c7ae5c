"""On-disk containers either side of the hot path.

.npz  -- exactly what the reference writes with np.savez_compressed
         (stylegan3_hvae_full.py:351-359, gumbel_softmax_compression.py:289-297).
.cabac -- two flavours:
   "packed"    (default) a WORKING container: header = pickle byte length, payload = MSB-first
               packed bits; round-trips through load.
   "reference" byte-for-byte what cabac_compression.py:555-561 writes: header = len(metadata dict)
               (= 6, defect D4) and one byte per bit (defect D1).  Kept for fidelity tests; the
               reference's own loader cannot read it back, and neither does ours.
"""
import pickle
import struct

import numpy as np


def write_cabac(filename, encoded_bytes, metadata, flavour="packed"):
    with open(filename, "wb") as f:
        blob = pickle.dumps(metadata)
        if flavour == "reference":
            f.write(struct.pack("I", len(metadata)))
        elif flavour == "packed":
            f.write(struct.pack("I", len(blob)))
        else:
            raise ValueError(flavour)
        f.write(blob)
        f.write(encoded_bytes)


def read_cabac(filename):
    """Reads a "packed" container (cabac_compression.py:577-583 with the header meaning fixed)."""
    with open(filename, "rb") as f:
        (n,) = struct.unpack("I", f.read(4))
        metadata = pickle.loads(f.read(n))
        payload = f.read()
    return payload, metadata


def write_latent_npz(filename, w, resolution, bits, orig_size, comp_size):
    np.savez_compressed(filename, w=w, resolution=resolution, bits=bits, orig_size=orig_size, comp_size=comp_size,
                        compression_ratio=orig_size / comp_size)


def write_codes_npz(filename, codes, n_embeddings, resolution, orig_size, comp_size):
    np.savez_compressed(filename, codes=codes, n_embeddings=n_embeddings, resolution=resolution, orig_size=orig_size,
                        comp_size=comp_size, compression_ratio=orig_size / comp_size)
