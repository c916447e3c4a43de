import os, sys, time, torch, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
import wmattack
dev = "cuda"
xs = torch.rand(1, 3, 64, 64, device=dev).requires_grad_(True); gs = torch.rand(1, 3, 64, 64, device=dev)
layers = {"blur": wmattack.GaussianBlur(), "dj": wmattack.DiffJPEG(True, 64, 64, 50), "resize": lambda t: wmattack.Resize()(t, resize_ratio=0.75)}
for name, m in layers.items():
    for _ in range(20): m(xs).backward(gs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300): y = m(xs)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    for _ in range(300): y = m(xs); y.backward(gs)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    torch.autograd.set_multithreading_enabled(False)
    for _ in range(300): y = m(xs); y.backward(gs)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    torch.autograd.set_multithreading_enabled(True)
    print(f"{name}: fwd {1e6*(t1-t0)/300:.1f} us, fwd+bwd {1e6*(t2-t1)/300:.1f} us, fwd+bwd single-thread autograd {1e6*(t3-t2)/300:.1f} us")
m = layers["blur"]
pr = cProfile.Profile(); pr.enable()
for _ in range(300): y = m(xs); y.backward(gs)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
