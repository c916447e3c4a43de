"""Drop-in for the DiffJPEG part of the reference's utils/JPEG.py (:97-540)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from wmattack.modules import (  # noqa: E402,F401
    DiffJPEG, compress_jpeg, decompress_jpeg, diff_round, quality_to_factor, round_only_at_0,
)
