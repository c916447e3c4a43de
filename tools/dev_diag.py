import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200")); sys.path.insert(0, ROOT)
import wmattack
from oracle import attack_oracle as O
def rnd(shape, seed): return torch.rand(shape, generator=torch.Generator().manual_seed(seed))
def stats(e): return f"max={e.max():.2e} f>1e-5={float((e>1e-5).float().mean()):.2e} f>2e-5={float((e>2e-5).float().mean()):.2e} f>5e-5={float((e>5e-5).float().mean()):.2e}"
for (b,h,w) in ((3,64,96),(2,128,128)):
  for q in (50,75,95):
    for mode in (0,1):
        x, g = rnd((b,3,h,w),100+h+q), rnd((b,3,h,w),200+w+q)
        xo = x.double().requires_grad_(True); yo = O.diffjpeg(xo,q,mode); yo.backward(g.double())
        x32 = x.clone().requires_grad_(True); y32 = O.diffjpeg(x32,q,mode); y32.backward(g)
        xx = x.cuda().requires_grad_(True); y = wmattack.DiffJPEG(True,h,w,q,mode)(xx); y.backward(g.cuda())
        print(f"{b}x{h}x{w} q{q} m{mode} |g|max={xo.grad.abs().max():.2f}\n   ours-vs-64: {stats((xx.grad.cpu().double()-xo.grad).abs())}\n   cpu32-vs-64: {stats((x32.grad.double()-xo.grad).abs())}")
from tests.golden_util import T
x = T("xs32")
for rn, mode in (("r0",0),("cubic",1)):
    m = wmattack.DiffJPEG(True,32,32,quality=50,rounding=mode)
    cy,ccb,ccr = m.compress(x.cuda())
    for got,key in ((cy,"coef_y"),(ccb,"coef_cb"),(ccr,"coef_cr")):
        ref = T(f"diffjpeg/q50/{rn}/xs32/{key}")
        d=(got.cpu()-ref).abs(); i=d.argmax()
        print(rn,key,"max",float(d.max()),"at val",float(ref.flatten()[i]), float(got.cpu().flatten()[i]))
