"""Import surface of the reference's ``quantization`` package (quantization/__init__.py:8)."""
from .vq import QuantizedResult, ResidualVectorQuantizer  # noqa: F401
