"""Drop-in for the reference's noise_layers/dropout.py."""
from wmattack.modules import MaskDropout as Dropout  # noqa: F401  (dropout.py:4)
