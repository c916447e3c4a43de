"""Drop-in for the reference's noise_layers/crop.py."""
from wmattack.modules import Crop, Cropout  # noqa: F401
from wmattack.modules import ElementDropout as Dropout  # noqa: F401  (crop.py:136)
