"""Tiny driver for ncu: DiffJPEG q50 fwd+bwd at BASELINE config-2 size (64x3x512x512)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
import wmattack
b, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 512, 512)))
x = torch.rand(b, 3, h, w, device="cuda", requires_grad=True)
g = torch.rand(b, 3, h, w, device="cuda")
m = wmattack.DiffJPEG(True, h, w, quality=50)
for _ in range(3):
    y = m(x)
    y.backward(g)
    x.grad = None
torch.cuda.synchronize()
print("ok")
