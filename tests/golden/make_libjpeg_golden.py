"""Generate tests/golden/libjpeg_golden.npz with Pillow (the codec the reference's JpegTest calls,
noise_layers/jpeg.py:33-34): seeded 8-bit frames and what `Image.open(save(...))` returns for them.

    python tests/golden/make_libjpeg_golden.py

Pillow 12.2.0 / libjpeg-turbo (API 6.2) in the build container.  Only inputs and OUTPUTS are stored.
"""
import io
import os

import numpy as np
from PIL import Image, features


def frames():
    rng = np.random.RandomState(1234)
    out = {}
    for name, (h, w) in {"a": (16, 16), "b": (33, 47), "c": (40, 56), "d": (17, 18)}.items():
        out[name + "_rand"] = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        yy, xx = np.mgrid[0:h, 0:w]
        base = np.kron(rng.rand(h // 8 + 2, w // 8 + 2, 3), np.ones((8, 8, 1)))[:h, :w]
        s = 0.6 * base + 0.4 * (np.sin(xx / 7.0 + yy / 5.0)[..., None] * 0.5 + 0.5)
        out[name + "_smooth"] = (s * 255).astype(np.uint8)
    return out


def roundtrip(rgb, q, s):
    buf = io.BytesIO()
    Image.fromarray(rgb).save(buf, format="JPEG", quality=q, subsampling=s)
    return np.array(Image.open(io.BytesIO(buf.getvalue())), dtype=np.uint8)


def main():
    gold = {}
    for name, rgb in frames().items():
        gold[f"in/{name}"] = rgb
        for s in (0, 1, 2):
            for q in (10, 50, 90, 100):
                gold[f"out/{name}/s{s}/q{q}"] = roundtrip(rgb, q, s)
    gold["meta"] = np.array([f"Pillow {Image.__version__ if hasattr(Image, '__version__') else ''}",
                             f"jpeg {features.version('jpg')} turbo={features.check_feature('libjpeg_turbo')}"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libjpeg_golden.npz")
    np.savez_compressed(path, **gold)
    print(path, os.path.getsize(path), "bytes,", len(gold), "arrays")


if __name__ == "__main__":
    main()
