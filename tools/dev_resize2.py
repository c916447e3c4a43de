import os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
from wmattack import functional as WF
torch.manual_seed(0)
B,H,W=64,512,512
x = torch.rand(B,3,H,W,device="cuda"); g = torch.rand(B,3,H,W,device="cuda")
def check(r, tag):
    mid=(int(r*H),int(r*W))
    xx=x.clone().requires_grad_(True); y=WF.resize_roundtrip(xx,mid,"bicubic"); y.backward(g)
    xr=x.clone().requires_grad_(True)
    m=F.interpolate(xr,size=mid,mode="bicubic"); pre=F.interpolate(m,size=(H,W),mode="bicubic"); yr=torch.clamp(pre,0,1); yr.backward(g)
    e=(xx.grad-xr.grad).abs()
    print(tag,r,"fwd",float((y.detach()-yr.detach()).abs().max()),"bwd max",float(e.max()),"n>1e-4",int((e>1e-4).sum()))
    if e.max()>1e-4:
        idx=(e>1e-4).nonzero()
        print("  first bad:",idx[:3].tolist(),"planes:",sorted(set((idx[:,0]*3+idx[:,1]).tolist()))[:12], "rows",int(idx[:,2].min()),int(idx[:,2].max()),"cols",int(idx[:,3].min()),int(idx[:,3].max()))
for r in (0.5,0.75,1.25,1.5,0.75,0.5):
    check(r,"A")
    # hammer: many fwd/bwd in a row like the timing loops
    mid=(int(r*H),int(r*W)); xx=x.clone().requires_grad_(True)
    for _ in range(10):
        y=WF.resize_roundtrip(xx,mid,"bicubic")
    for _ in range(10):
        xx.grad=None; y.backward(g,retain_graph=True)
    check(r,"B")
