"""Drop-in for the reference's noise_layers/gaussian_noise.py."""
from wmattack.modules import GN  # noqa: F401
