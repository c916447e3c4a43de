"""Drop-in for the reference's noise_layers/combined.py."""
from wmattack.modules import Combined  # noqa: F401
