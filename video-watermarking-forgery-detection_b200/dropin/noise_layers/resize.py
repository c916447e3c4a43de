"""Drop-in for the reference's noise_layers/resize.py."""
from wmattack.modules import Resize, random_float  # noqa: F401
