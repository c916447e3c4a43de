"""Drop-in for the reference's noise_layers/gaussian_filter.py."""
from wmattack.modules import GF  # noqa: F401
