"""Host-side cost of one layer call on a tiny batch (developer tool, GPU box): wall time per eager
forward / forward+backward and a cProfile of where the Python time goes."""
import cProfile, os, pstats, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
import wmattack

dev = "cuda"
x = torch.rand(1, 3, 256, 256, device=dev)
g = torch.rand(1, 3, 256, 256, device=dev)
rs = wmattack.Resize()
layers = {"diffjpeg": wmattack.DiffJPEG(True, 256, 256, quality=50), "jpegcompression": wmattack.JpegCompression(dev),
          "blur": wmattack.GaussianBlur(), "median3": wmattack.MiddleBlur(3), "gaussian": wmattack.Gaussian(),
          "resize": lambda t: rs(t, resize_ratio=0.75)}


def wall(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


for name, layer in layers.items():
    with torch.no_grad():
        tf = wall(lambda: layer(x))
    xr = x.clone().requires_grad_(True)
    def step():
        (gx,) = torch.autograd.grad(layer(xr), xr, g)
    print(f"{name:16s} forward (no grad) {tf:6.1f} us/call   forward+backward {wall(step):6.1f} us/step", flush=True)

# the same measurement for a built-in torch op: the floor autograd itself imposes on this host
xr = x.clone().requires_grad_(True)
with torch.no_grad():
    tf = wall(lambda: x * 1.5)
print(f"{'torch x * 1.5':16s} forward (no grad) {tf:6.1f} us/call   forward+backward {wall(lambda: torch.autograd.grad(xr * 1.5, xr, g)):6.1f} us/step", flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "blur"
layer = layers[which]
pr = cProfile.Profile()
xr = x.clone().requires_grad_(True)
pr.enable()
for _ in range(3000):
    torch.autograd.grad(layer(xr), xr, g)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
