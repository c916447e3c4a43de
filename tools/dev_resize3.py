import os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
from wmattack import functional as WF
torch.manual_seed(0)
for (B,H,W,r) in [(1,20,28,0.5),(1,64,128,0.5),(1,128,256,0.5),(1,512,512,0.5),(1,64,128,0.55),(1,64,128,0.47)]:
    x = torch.rand(B,1,H,W,device="cuda")
    mid=(int(r*H),int(r*W))
    y=WF.resize_roundtrip(x,mid,"bicubic")
    m=F.interpolate(x,size=mid,mode="bicubic"); yr=torch.clamp(F.interpolate(m,size=(H,W),mode="bicubic"),0,1)
    e=(y-yr).abs()
    t=WF._RESIZE_TABLES[(str(x.device),H,W,mid[0],mid[1],1)]
    print((B,H,W,r),"fwd max",float(e.max()),"n>1e-4",int((e>1e-4).sum()),"overflow",int(t[-4:].view(torch.int32)[0]), "tables", t.numel())
    if e.max()>1e-4:
        idx=(e>1e-4).nonzero()
        print("   rows",sorted(set(idx[:,2].tolist()))[:20],"cols",sorted(set(idx[:,3].tolist()))[:40])
