"""Drop-in for the reference's noise_layers/__init__.py (exports of :4-19)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from wmattack.modules import get_random_float, get_random_int  # noqa: E402,F401
from .identity import Identity  # noqa: E402,F401
from .crop import Crop, Cropout, Dropout  # noqa: E402,F401
from .gaussian_noise import GN  # noqa: E402,F401
from .middle_filter import MiddleBlur  # noqa: E402,F401
from .gaussian_filter import GF  # noqa: E402,F401
from .salt_pepper_noise import SaltPepper  # noqa: E402,F401
from .jpeg import Jpeg, JpegSS, JpegMask, JpegTest  # noqa: E402,F401
from .combined import Combined  # noqa: E402,F401
