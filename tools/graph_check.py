"""CUDA-graph capture of attack-layer calls (developer check, GPU box only): every launch goes to
torch's current stream and the library never allocates, so a forward+backward can be captured
once and replayed on new data; prints eager vs replay time per step (small batches are launch-bound).

    python tools/graph_check.py [B H W]
"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
import wmattack

dev = "cuda"
B, H, W = (int(v) for v in (sys.argv[1:4] or (1, 256, 256)))
rs = wmattack.Resize()
layers = {"diffjpeg": wmattack.DiffJPEG(True, H, W, quality=50), "jpegcompression": wmattack.JpegCompression(dev),
          "blur": wmattack.GaussianBlur(), "median3": wmattack.MiddleBlur(3),
          "resize": lambda t: rs(t, resize_ratio=0.75), "jpegss": wmattack.JpegSS(50)}


def timeit(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


for name, layer in layers.items():
    xs = torch.rand(B, 3, H, W, device=dev, requires_grad=True)
    gs = torch.rand(B, 3, H, W, device=dev)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):                                   # warm-up off the capturing stream
        for _ in range(3):
            (gx,) = torch.autograd.grad(layer(xs), xs, gs)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y_cap = layer(xs)
        (gx_cap,) = torch.autograd.grad(y_cap, xs, gs)
    x2, g2 = torch.rand(B, 3, H, W, device=dev), torch.rand(B, 3, H, W, device=dev)
    with torch.no_grad():
        xs.copy_(x2); gs.copy_(g2)
    graph.replay(); torch.cuda.synchronize()
    xe = x2.clone().requires_grad_(True)
    ye = layer(xe)
    (ge,) = torch.autograd.grad(ye, xe, g2)
    ok = torch.equal(ye, y_cap) and torch.equal(ge, gx_cap)
    t_eager = timeit(lambda: torch.autograd.grad(layer(xs), xs, gs))
    t_graph = timeit(graph.replay)
    print(f"{name:16s} replay==eager: {ok}   eager {t_eager:7.1f} us/step   graph replay {t_graph:7.1f} us/step", flush=True)
