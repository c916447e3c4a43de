"""Drop-in for the reference's noise_layers/gaussian_blur.py."""
from wmattack.modules import GaussianBlur  # noqa: F401
