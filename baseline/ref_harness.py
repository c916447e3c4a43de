"""Import harness for the UNMODIFIED reference files vendored under baseline/_ref/ (git-ignored,
written by baseline/vendor_reference.py).  Used ONLY by bench.py's reference legs; the product
(wmattack/) never imports this.

The reference stays byte-identical; the harness supplies what its imports need and this image lacks
(SURVEY.md Appendix C):
  1. empty stand-ins for `matplotlib` / `matplotlib.pyplot` (imported, never called on the path);
  2. a stand-in `kornia.filters` with kornia 0.6.x semantics for MedianBlur / GaussianBlur2d
     (kornia is a third-party dependency of the reference that is neither vendored, pinned nor
     installed: the reference's MiddleBlur / GF delegate to it);
  3. CPU runs only: `.cuda()` made a no-op, because DiffJPEG.__init__, GaussianBlur.forward and
     Gaussian.forward hard-code `.cuda()` (utils/JPEG.py:532-533, gaussian_blur.py:57, gaussian.py:13).
     On the GPU box (device='cuda') shim 3 is not installed and the same files run eagerly on the B200.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "utils", "JPEG.py")) and os.path.exists(os.path.join(REF, "noise_layers", "__init__.py"))


class _MedianBlur(nn.Module):
    """kornia 0.6.x MedianBlur: zero-pad, one-hot conv to [B*C, k*k, H, W], torch.median over dim."""

    def __init__(self, kernel_size):
        super().__init__()
        self.k = kernel_size

    def forward(self, x):
        kh, kw = self.k
        b, c, h, w = x.shape
        kernel = torch.eye(kh * kw, dtype=x.dtype, device=x.device).view(kh * kw, 1, kh, kw)
        feat = F.conv2d(x.reshape(b * c, 1, h, w), kernel, padding=((kh - 1) // 2, (kw - 1) // 2), stride=1)
        return feat.view(b, c, -1, h, w).median(dim=2)[0]


class _GaussianBlur2d(nn.Module):
    def __init__(self, kernel_size, sigma, border_type="reflect"):
        super().__init__()
        self.k, self.s, self.border = kernel_size, sigma, border_type

    def forward(self, x):
        def taps(k, s):
            xs = torch.arange(k, dtype=x.dtype, device=x.device) - k // 2
            g = torch.exp(-xs ** 2 / (2 * s ** 2))
            return g / g.sum()
        ky, kx = taps(self.k[0], self.s[0]), taps(self.k[1], self.s[1])
        c = x.shape[1]
        w2 = torch.outer(ky, kx).view(1, 1, *self.k).repeat(c, 1, 1, 1)
        ry, rx = self.k[0] // 2, self.k[1] // 2
        return F.conv2d(F.pad(x, (rx, rx, ry, ry), mode=self.border), w2, groups=c)




_installed = {"done": False}


@contextlib.contextmanager
def cpu_mode():
    """Shim 3: while active, `.cuda()` on modules and tensors is a no-op (CPU runs of files that
    hard-code `.cuda()`); restored on exit so the same process can run the GPU-eager leg afterwards."""
    saved = (nn.Module.cuda, torch.Tensor.cuda)
    nn.Module.cuda = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        nn.Module.cuda, torch.Tensor.cuda = saved


def install() -> None:
    """Make `from utils.JPEG import DiffJPEG` / `import noise_layers` resolve to baseline/_ref."""
    if not available():
        raise FileNotFoundError(f"{REF} is empty: run `python baseline/vendor_reference.py` in the build container "
                                "(needs /root/reference); the directory ships to the GPU box with the snapshot")
    if not _installed["done"]:
        for name in ("matplotlib", "matplotlib.pyplot"):
            if name not in sys.modules:
                try:
                    __import__(name)
                except Exception:
                    sys.modules[name] = types.ModuleType(name)
        if not hasattr(sys.modules["matplotlib"], "pyplot"):
            sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        try:
            import kornia.filters  # noqa: F401  (a real kornia wins if one is ever installed)
        except Exception:
            kornia = types.ModuleType("kornia")
            kf = types.ModuleType("kornia.filters")
            kf.MedianBlur, kf.GaussianBlur2d = _MedianBlur, _GaussianBlur2d
            kornia.filters = kf
            sys.modules["kornia"], sys.modules["kornia.filters"] = kornia, kf
        # the repo's own `utils` / `noise_layers` names must not shadow the reference's
        for name in [n for n in sys.modules if n == "utils" or n.startswith("utils.") or n == "noise_layers" or n.startswith("noise_layers.")]:
            del sys.modules[name]
        sys.path.insert(0, REF)
        _installed["done"] = True


def kornia_is_stub() -> bool:
    return getattr(sys.modules.get("kornia.filters"), "MedianBlur", None) is _MedianBlur


def _by_path(name: str, rel: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def layers(device: str = "cpu"):
    """Namespace of the reference's own classes (imported from baseline/_ref).  For CPU runs wrap
    construction AND calls in `with cpu_mode():`."""
    install()
    from utils.JPEG import DiffJPEG
    import noise_layers as NL
    from noise_layers.jpeg_compression import JpegCompression
    from noise_layers.gaussian_blur import GaussianBlur
    from noise_layers.gaussian import Gaussian
    from noise_layers.middle_filter import MiddleBlur
    from noise_layers.resize import Resize
    from noise_layers.combined import Combined
    ns = types.SimpleNamespace(DiffJPEG=DiffJPEG, JpegCompression=JpegCompression, GaussianBlur=GaussianBlur,
                               Gaussian=Gaussian, MiddleBlur=MiddleBlur, Resize=Resize, Combined=Combined,
                               Jpeg=NL.Jpeg, JpegSS=NL.JpegSS, JpegMask=NL.JpegMask, Identity=NL.Identity,
                               Crop=NL.Crop, SaltPepper=NL.SaltPepper, GN=NL.GN)
    ns.Quantization = _by_path("ref_Quantization", "models/modules/Quantization.py").Quantization
    return ns


def networks():
    """config 4's encoder / localiser classes (models/invertible_net.py, network/UNet.py), loaded by
    path so that the reference's models/__init__.py (which imports the whole trainer) is not run."""
    install()
    inv = _by_path("ref_invertible_net", "models/invertible_net.py")
    unet = _by_path("ref_unet", "network/UNet.py")
    return inv, unet
