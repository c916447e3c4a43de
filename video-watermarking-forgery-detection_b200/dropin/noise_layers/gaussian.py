"""Drop-in for the reference's noise_layers/gaussian.py."""
from wmattack.modules import Gaussian  # noqa: F401
