"""B200-native residual vector quantizer for EnCodec (drop-in for the reference's
``quantization`` package; hot path in sm_100a CUDA behind ``include/rvq_b200.h``)."""
from .quantization import QuantizedResult, ResidualVectorQuantizer  # noqa: F401

__all__ = ["QuantizedResult", "ResidualVectorQuantizer"]
