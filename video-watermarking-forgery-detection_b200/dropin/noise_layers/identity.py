"""Drop-in for the reference's noise_layers/identity.py."""
from wmattack.modules import Identity  # noqa: F401
