"""Inference-time vector quantisation with the reference's surface
(modules/quantization/{vq,core_vq}.py)."""
from .vq import ResidualVectorQuantizer  # noqa: F401
