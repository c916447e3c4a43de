"""Developer smoke check of the DiffJPEG kernels against the oracle (GPU box)."""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import attack_oracle as O
lib = C.CDLL(os.path.join(ROOT, "video-watermarking-forgery-detection_b200", "wmattack", "libwmattack.so"))
lib.wm_last_error.restype = C.c_char_p
vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
lib.wm_diffjpeg_fwd.argtypes = [vp, i64, i64, i64, vp, i32, i32, i32, f32, vp, i32, vp]
lib.wm_diffjpeg_bwd.argtypes = [vp, i64, i64, i64, vp, i64, i64, i64, vp, i32, i32, i32, f32, vp, i32, vp]
lib.wm_diffjpeg_compress.argtypes = [vp, i64, i64, i64, vp, vp, vp, i32, i32, i32, f32, vp, i32, vp]
lib.wm_diffjpeg_decompress.argtypes = [vp, vp, vp, vp, i32, i32, i32, f32, vp, vp]

def chk(rc):
    if rc: raise RuntimeError(f"rc={rc}: {lib.wm_last_error().decode()}")

def fwd(x, factor, mode):
    B, _, H, W = x.shape; y = torch.empty_like(x)
    chk(lib.wm_diffjpeg_fwd(x.data_ptr(), 3*H*W, H*W, W, y.data_ptr(), B, H, W, factor, None, mode, torch.cuda.current_stream().cuda_stream))
    return y
def bwd(x, g, factor, mode):
    B, _, H, W = x.shape; gx = torch.empty_like(x)
    chk(lib.wm_diffjpeg_bwd(x.data_ptr(), 3*H*W, H*W, W, g.data_ptr(), 3*H*W, H*W, W, gx.data_ptr(), B, H, W, factor, None, mode, torch.cuda.current_stream().cuda_stream))
    return gx

dev = "cuda"
for (B, H, W) in ((2, 32, 32), (3, 64, 96), (1, 16, 16), (5, 48, 272)):
    for q in (50, 75, 10, 95):
        for mode in (0, 1, 2, 3):
            g = torch.Generator().manual_seed(B*H+q)
            x = torch.rand(B, 3, H, W, generator=g); gy = torch.rand(B, 3, H, W, generator=g)
            fac = O.quality_to_factor(q)
            xx = x.double().requires_grad_(True)
            yo = O.diffjpeg(xx, q, mode); yo.backward(gy.double())
            y = fwd(x.to(dev), fac, mode).cpu(); gx = bwd(x.to(dev), gy.to(dev), fac, mode).cpu()
            dy = (y.double()-yo.detach()).abs().max().item(); dg = (gx.double()-xx.grad).abs().max().item()
            flag = "" if (dy < 1e-5 and dg < 1e-5) or mode == 2 else "  <<<<"
            print(f"B{B} {H}x{W} q{q} mode{mode}: dy={dy:.2e} dgx={dg:.2e} |gx|max={xx.grad.abs().max():.2f}{flag}")
# compress / decompress
x = torch.rand(2, 3, 32, 48, generator=torch.Generator().manual_seed(1))
x8 = torch.round(x*255)/255
for mode in (0, 2):
    yo, cbo, cro = O.diffjpeg_compress(x8.double(), 1.0, mode)
    xd = x8.to(dev)
    cy = torch.empty(2, 24, 8, 8, device=dev); ccb = torch.empty(2, 6, 8, 8, device=dev); ccr = torch.empty_like(ccb)
    chk(lib.wm_diffjpeg_compress(xd.data_ptr(), 3*32*48, 32*48, 48, cy.data_ptr(), ccb.data_ptr(), ccr.data_ptr(), 2, 32, 48, 1.0, None, mode, None))
    torch.cuda.synchronize()
    for a, b, n in ((cy, yo, "y"), (ccb, cbo, "cb"), (ccr, cro, "cr")):
        d = (a.cpu().double()-b).abs()
        print(f"compress mode{mode} {n}: max={d.max():.3e} mismatches(>0.5)={(d>0.5).sum().item()}")
    out = torch.empty(2, 3, 32, 48, device=dev)
    chk(lib.wm_diffjpeg_decompress(cy.data_ptr(), ccb.data_ptr(), ccr.data_ptr(), out.data_ptr(), 2, 32, 48, 1.0, None, None))
    ref = O.diffjpeg_decompress(cy.cpu().double(), ccb.cpu().double(), ccr.cpu().double(), 32, 48, 1.0)
    print(f"decompress mode{mode}: {(out.cpu().double()-ref).abs().max():.3e}")

# timing at config-2 size
B, H, W = 64, 512, 512
x = torch.rand(B, 3, H, W, device=dev); gy = torch.rand(B, 3, H, W, device=dev)
for name, fn, bpp in (("fwd", lambda: fwd(x, 1.0, 0), 24), ("bwd", lambda: bwd(x, gy, 1.0, 0), 36)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/20
    px = B*H*W
    print(f"{name}: {ms*1e3:.1f} us  {px/ms/1e3:.0f} Mpix/s  {px*bpp/ms/1e6:.0f} GB/s")
