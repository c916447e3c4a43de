"""Drop-in for the reference's noise_layers/jpeg_compression.py."""
from wmattack.modules import JpegCompression  # noqa: F401
