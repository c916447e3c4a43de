"""Drop-in for the reference's noise_layers/jpeg.py."""
from wmattack.modules import Jpeg, JpegBasic, JpegMask, JpegSS, JpegTest  # noqa: F401
