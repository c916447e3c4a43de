"""CUDA-graph capture of attack-layer calls (developer check, GPU box only): every launch goes to
torch's current stream and the library never allocates, so a forward+backward can be captured
once and replayed; prints eager vs replay time per step for a small (launch-bound) batch."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
import wmattack

dev = "cuda"
B, H, W = (int(v) for v in (sys.argv[1:4] or (1, 256, 256)))
layers = {"diffjpeg": wmattack.DiffJPEG(True, H, W, quality=50), "jpegcompression": wmattack.JpegCompression(dev),
          "blur": wmattack.GaussianBlur(), "median3": wmattack.MiddleBlur(3), "resize": None, "jpegss": wmattack.JpegSS(50)}
rs = wmattack.Resize()
layers["resize"] = lambda t: rs(t, resize_ratio=0.75)

def timeit(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6

for name, layer in layers.items():
    xs = torch.rand(B, 3, H, W, device=dev, requires_grad=True)
    gs = torch.rand(B, 3, H, W, device=dev)
    # eager reference
    y = layer(xs); y.backward(gs); y_ref, g_ref = y.detach().clone(), xs.grad.detach().clone(); xs.grad = None
    # warm-up on a side stream, then capture forward + backward in one graph
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            y = layer(xs); y.backward(gs); xs.grad = None
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y_cap = layer(xs)
        (gx_cap,) = torch.autograd.grad(y_cap, xs, gs)
    # replay on new data written into the static buffers
    x2 = torch.rand(B, 3, H, W, device=dev); g2 = torch.rand(B, 3, H, W, device=dev)
    with torch.no_grad():
        xs.copy_(x2); gs.copy_(g2)
    graph.replay(); torch.cuda.synchronize()
    xe = x2.clone().requires_grad_(True); ye = layer(xe); ye.backward(g2)
    ok = torch.equal(ye, y_cap) and torch.equal(xe.grad, gx_cap)
    def eager():
        xs.grad = None
        yy = layer(xs); yy.backward(gs)
    print(f"{name:16s} capture ok={ok}  eager {timeit(eager):7.1f} us/step   graph replay {timeit(graph.replay):7.1f} us/step", flush=True)
