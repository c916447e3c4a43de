import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))
from wmattack import functional as WF
import torch.nn.functional as F
tp=[0.3192,0.3616,0.3192]
for shape in [(2,3,20,28),(1,3,64,64),(1,3,66,136),(1,3,128,256),(2,3,70,300),(1,1,8,4),(6,1,20,28),(1,3,20,128), (1,3,20,132),(1,3,20,136),(1,3,20,140)]:
    x=torch.rand(*shape,device="cuda")
    try:
        y=WF.gaussian_blur(x,tp,0)
        w2=torch.tensor(tp,device="cuda"); w2=(w2[:,None]*w2[None,:]).expand(shape[1],1,3,3).contiguous()
        print(shape,"ok",float((y-F.conv2d(x,w2,padding=1,groups=shape[1])).abs().max()))
    except Exception as e:
        print(shape,"FAIL",str(e)[-80:])
