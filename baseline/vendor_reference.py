#!/usr/bin/env python
"""Copy the reference's hot-path files, UNMODIFIED, into the git-ignored baseline/_ref/.

    python baseline/vendor_reference.py [--ref /root/reference]

Why: the reference (yingqichao/video-watermarking-forgery-detection) is pure Python with no
setup.py / pyproject.toml, so `pip install --target baseline/_ref` has nothing to install; and
/root/reference does not exist on the GPU box.  baseline/_ref/ is listed in .gitignore (never
enters history) but NOT in .gpurunignore, so these files travel with the repo snapshot exactly like
the built libwmattack.so does.  bench.py's reference arm (`--impl reference`, `cpu_baseline`,
`reference_gpu_eager`) imports them through baseline/ref_harness.py; nothing under wmattack/ does.

Files = SURVEY 8(a)'s hot path (noise_layers/*, utils/JPEG*.py, compression/decompression,
Quantization) plus the two network definitions BASELINE config 4 names (invertible_net, UNet).
A MANIFEST with sha256 of every file is written so a reader can verify they are byte-identical.
"""
import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")

FILES = [
    "noise_layers/__init__.py", "noise_layers/combined.py", "noise_layers/crop.py", "noise_layers/dropout.py",
    "noise_layers/gaussian.py", "noise_layers/gaussian_blur.py", "noise_layers/gaussian_filter.py",
    "noise_layers/gaussian_noise.py", "noise_layers/identity.py", "noise_layers/jpeg.py",
    "noise_layers/jpeg_compression.py", "noise_layers/middle_filter.py", "noise_layers/resize.py",
    "noise_layers/salt_pepper_noise.py",
    "utils/__init__.py", "utils/JPEG.py", "utils/JPEG_utils.py", "utils/compression.py", "utils/decompression.py",
    "models/modules/Quantization.py",
    "models/invertible_net.py", "network/UNet.py",          # config 4's encoder / localiser definitions
]


def vendor(ref: str = "/root/reference", dest: str = DEST) -> dict:
    if not os.path.isdir(ref):
        raise FileNotFoundError(f"reference tree {ref} not present (expected only in the build container)")
    manifest = {}
    for rel in FILES:
        src = os.path.join(ref, rel)
        dst = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": "yingqichao/video-watermarking-forgery-detection (unmodified copies)", "sha256": manifest}, f, indent=1)
    return manifest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    m = vendor(a.ref)
    print(f"vendored {len(m)} files into {DEST}")
