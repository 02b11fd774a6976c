import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from ptbxl_multimodal_b200._lib import lib, check, ptr, stream
DEV='cuda:0'; BF=torch.bfloat16
def gen(*s, seed=0): return torch.randn(*s, generator=torch.Generator().manual_seed(seed))
def to_blocked(x):
    b,c,l=x.shape; return x.reshape(b,c//8,8,l).permute(0,1,3,2).contiguous().to(BF)
for (B,Ci,Co,L) in [(1,32,64,40),(1,32,64,128),(1,32,128,128),(2,128,256,625),(6,128,256,125),(1,128,256,125)]:
    Cip=(Ci+15)//16*16
    x=torch.zeros(B,Cip,L); x[:,:Ci]=gen(B,Ci,L,seed=6); dy=gen(B,Co,L,seed=7)
    xr=x.to(BF).float()[:,:Ci]; dyr=dy.to(BF).float()
    w=torch.zeros(Co,Ci,15,requires_grad=True); F.conv1d(xr,w,None,padding=7).backward(dyr)
    ws=torch.empty(lib.ecgb200_conv1d_wgrad_bf16_ws_bytes(B,Ci,Co,L),dtype=torch.uint8,device=DEV)
    dw=torch.full((Co,Ci,15),float('nan'),device=DEV); db=torch.empty(Co,device=DEV)
    check(lib.ecgb200_conv1d_wgrad_bf16(ptr(to_blocked(dy).to(DEV)),ptr(to_blocked(x).to(DEV)),ptr(dw),ptr(db),None,0,ptr(ws),B,Ci,Co,L,stream()),'wg')
    torch.cuda.synchronize()
    err=(dw.cpu()-w.grad).abs(); bad=err>1e-2*w.grad.abs().max()
    print((B,Ci,Co,L),'bad frac',float(bad.float().mean()))
    if bad.any():
        print('  bad o:',sorted(set(bad.nonzero()[:,0].tolist()))[:40])
        print('  bad c:',sorted(set(bad.nonzero()[:,1].tolist()))[:40])
        print('  bad k:',sorted(set(bad.nonzero()[:,2].tolist())))
        o,c=bad.nonzero()[0,:2].tolist()
        print('  sample (o,c)=',o,c,'mine',dw[o,c].cpu().numpy().round(2),'ref',w.grad[o,c].numpy().round(2))
        # hypotheses
        xx=xr[0]; dd=dyr[0]
        def corr(a,bv):
            ap=F.pad(bv,(7,7)); return torch.stack([(a*ap[k:k+L]).sum() for k in range(15)])
        for name,a,bv in [('dy[o]*dy[c]',dd[o],dd[c] if c<Co else dd[0]),('x[o]*x[c]',xx[o] if o<Ci else xx[0],xx[c]),('dy[o]*x[c]',dd[o],xx[c])]:
            print('   ',name,corr(a,bv).numpy().round(2))
