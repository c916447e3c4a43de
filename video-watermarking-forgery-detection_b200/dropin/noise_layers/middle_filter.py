"""Drop-in for the reference's noise_layers/middle_filter.py."""
from wmattack.modules import MiddleBlur  # noqa: F401
