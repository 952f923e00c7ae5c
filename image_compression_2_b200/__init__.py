"""image_compression_2_b200 -- B200 (sm_100a) implementation of the latent compression hot path of
yubster4525/image_compression_2: W+ quantisation, context-adaptive arithmetic coding, decoding and
dequantisation, behind the reference's own Python API.  See DESIGN.md / INTEGRATION.md.

Importing the package does not need a GPU; every operation does (there is no CPU fallback).
"""
from . import codec, coder, compressors, containers, pipeline, sharding, stats  # noqa: F401
from .coder import ContextModel, cabac_decode, cabac_encode  # noqa: F401
from .compressors import (CABACCompressor, GumbelSoftmaxCompressor, GumbelSoftmaxDiscretization,  # noqa: F401
                          StyleGAN3Compressor)
from .pipeline import LatentPipeline  # noqa: F401

__version__ = "0.1.0"
