import os, sys, torch
sys.path.insert(0, "video-watermarking-forgery-detection_b200")
import wmattack
dev = "cuda"
torch.manual_seed(0)
for dt in (torch.bfloat16, torch.float16):
    for shape in ((2, 3, 80, 136), (1, 3, 37, 264), (1, 3, 130, 8)):
        x = (torch.rand(*shape, device=dev) * 1.2 - 0.1).to(dt)
        g = torch.rand(*shape, device=dev)
        for layer, kw in [(wmattack.GaussianBlur(3), {}), (wmattack.GaussianBlur(7), {}), (wmattack.MiddleBlur(3), {}), (wmattack.MiddleBlur(5), {}),
                          (wmattack.Resize(), {"resize_ratio": 0.75}), (wmattack.Resize(), {"resize_ratio": 1.5}),
                          (wmattack.Resize(interpolation_method="bilinear"), {"resize_ratio": 0.5})]:
            name = f"{dt} {shape} {type(layer).__name__} {kw}"
            try:
                xa = x.clone().requires_grad_(True)
                ya = layer(xa, **kw); torch.cuda.synchronize()
                print("fwd ok ", name, flush=True)
                ya.backward(g); torch.cuda.synchronize()
                print("bwd ok ", name, flush=True)
                xb = x.float().requires_grad_(True)
                yb = layer(xb, **kw); yb.backward(g); torch.cuda.synchronize()
                print("   y equal", torch.equal(ya, yb), " gx equal", torch.equal(xa.grad, xb.grad.to(dt)), xa.grad.dtype,
                      " max|dy|", float((ya - yb).abs().max()), " max|dg|", float((xa.grad.float() - xb.grad).abs().max()), flush=True)
            except Exception as e:
                print("FAIL", name, str(e)[:300], flush=True)
                sys.exit(1)
