import os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
from wmattack import functional as WF
torch.manual_seed(0)
for (B,H,W,r) in [(2,512,512,0.75),(64,512,512,0.75),(2,512,512,1.5),(8,512,512,1.5),(1,512,512,0.75),(2,256,384,0.75)]:
    x = torch.rand(B,3,H,W,device="cuda"); g = torch.rand(B,3,H,W,device="cuda")
    mid=(int(r*H),int(r*W))
    xx=x.clone().requires_grad_(True); y=WF.resize_roundtrip(xx,mid,"bicubic"); y.backward(g)
    xr=x.clone().requires_grad_(True)
    m=F.interpolate(xr,size=mid,mode="bicubic"); pre=F.interpolate(m,size=(H,W),mode="bicubic"); yr=torch.clamp(pre,0,1); yr.backward(g)
    e=(xx.grad-xr.grad).abs()
    print((B,H,W,r),"fwd",float((y-yr).abs().max()),"bwd max",float(e.max()),"n>1e-4",int((e>1e-4).sum()))
    if e.max()>1e-4:
        idx=(e>1e-4).nonzero()
        print("  first bad:",idx[:5].tolist(),"planes:",sorted(set((idx[:,0]*3+idx[:,1]).tolist()))[:10], "rows range",int(idx[:,2].min()),int(idx[:,2].max()),"cols",int(idx[:,3].min()),int(idx[:,3].max()))
        # is it a mask flip? compare with gradient computed using OUR forward's mask decision
        near=((pre.detach()).abs()<2e-6)|((pre.detach()-1).abs()<2e-6)
        print("  outputs within 2e-6 of a clamp bound:",int(near.sum()))
