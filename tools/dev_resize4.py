import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200")); sys.path.insert(0, ROOT)
from wmattack import functional as WF
from oracle import attack_oracle as O
Fi = torch.nn.functional.interpolate
def rnd(shape, seed): return torch.rand(shape, generator=torch.Generator().manual_seed(seed))
h,w,r,mode=130,260,0.66,"bicubic"
x, g = rnd((2,2,h,w),43), rnd((2,2,h,w),143)
mid=(int(r*h),int(r*w))
xx=x.cuda().requires_grad_(True); y=WF.resize_roundtrip(xx,mid,mode); y.backward(g.cuda())
xo=x.clone().requires_grad_(True); pre=Fi(Fi(xo,size=list(mid),mode=mode),size=[h,w],mode=mode); pre.clamp(0,1).backward(g)
x6=x.double().requires_grad_(True); pre6=O.interpolate(O.interpolate(x6,mid,mode),(h,w),mode); pre6.clamp(0,1).backward(g.double())
e=(xx.grad.cpu()-xo.grad).abs(); e6=(xx.grad.cpu().double()-x6.grad).abs(); ec=(xo.grad.double()-x6.grad).abs()
print("ours-vs-cpu32",float(e.max()),"ours-vs-fp64",float(e6.max()),"cpu32-vs-fp64",float(ec.max()))
idx=(e>1e-5).nonzero(); print(idx[:10].tolist(), len(idx))
d=(pre.detach()-pre6.detach()).abs(); print("pre cpu32 vs fp64",float(d.max()))
near=((pre6.detach().abs()<3e-5)|((pre6.detach()-1).abs()<3e-5)).nonzero(); print("near-bound outputs:",near[:10].tolist(), [float(pre6.detach()[tuple(i)]) for i in near[:5]])
