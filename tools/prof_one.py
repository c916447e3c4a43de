"""Tiny driver for ncu: runs ONE op a few times at BASELINE config-2 size.
    python tools/prof_one.py resize|blur|median|diffjpeg|jpeg8 [B H W]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
import wmattack
from wmattack import functional as WF
op = sys.argv[1]
b, h, w = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (64, 512, 512)))
x = torch.rand(b, 3, h, w, device="cuda", requires_grad=True)
g = torch.rand(b, 3, h, w, device="cuda")
fns = {
    "resize": lambda t: WF.resize_roundtrip(t, (int(0.75 * h), int(0.75 * w)), "bicubic"),
    "resize15": lambda t: WF.resize_roundtrip(t, (int(1.5 * h), int(1.5 * w)), "bicubic"),
    "blur": lambda t: WF.gaussian_blur(t, [0.3192, 0.3616, 0.3192], 0),
    "median": lambda t: WF.median_blur(t, 3),
    "median5": lambda t: WF.median_blur(t, 5),
    "diffjpeg": wmattack.DiffJPEG(True, h, w, quality=50),
    "jpeg8": wmattack.JpegCompression("cuda"),
    "jpegss": wmattack.JpegSS(50),
    "noise": wmattack.Gaussian(),
    "crop": lambda t: wmattack.Crop()(t, apex=(int(0.2 * h), int(0.2 * h) + int(0.7 * h), int(0.1 * w), int(0.1 * w) + int(0.75 * w)))[0],
    "codec": lambda t: WF.jpeg_codec(t * 2 - 1, 75, 2, "signed"),
}
f = fns[op]
for _ in range(3):
    y = f(x)
    if y.requires_grad:
        y.backward(g)
        x.grad = None
torch.cuda.synchronize()
print("ok")
