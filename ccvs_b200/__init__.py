"""ccvs_b200 — B200-native (sm_100a) implementation of the CCVS latent vector-quantization path.

Public surface (mirrors the reference's models/skip_vid_generator/modules/quantize.py):
    VectorQuantizer   drop-in nn.Module (forward / embed_code / embedding.weight)
    EncoderTail       the encoder's last 1x1 ConvLayer (+ output normalisation) that produces the latents
    ops               tensor-level wrappers over the C ABI in include/ccvsq.h
    dist              frame sharding + packed all-reduce helpers for multi-GPU training statistics

Importing `VectorQuantizer` / `ops` loads ccvs_b200/lib/libccvsq.so and raises if it is missing:
there is no CPU or pure-PyTorch fallback.
"""
from ._lib import Layout, build, load  # noqa: F401

__all__ = ["VectorQuantizer", "EncoderTail", "ops", "dist", "Layout", "build", "load"]


def __getattr__(name):
    # lazy so that `import ccvs_b200; ccvs_b200.build()` works before the .so exists
    if name == "VectorQuantizer":
        from .quantize import VectorQuantizer
        return VectorQuantizer
    if name == "EncoderTail":
        from .encoder_tail import EncoderTail
        return EncoderTail
    if name in ("ops", "dist", "quantize", "encoder_tail"):
        import importlib
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
