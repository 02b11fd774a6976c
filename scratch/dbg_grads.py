import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import torch.nn.functional as F
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200 import functional as Fn
from oracle import ecg_oracle as O
DEV='cuda:0'
def rel(a,b):
    a=a.detach().double().cpu(); b=b.detach().double().cpu()
    return float((a-b).abs().max()/b.abs().max().clamp_min(1e-30))
for (B,T) in [(6,1000),(3,250),(16,1000)]:
    x,y=O.synth_batch(B,T,5,seed=0)
    sd=O.init_state_dict('cnn',5,seed=42)
    # reference with intermediate grads
    keys=O.param_keys(sd)
    leaves={k:sd[k].clone().requires_grad_(True) for k in keys}
    work=dict(sd); work.update(leaves)
    h=x; acts=[]
    for i in range(4):
        p,a=O.conv_block(work,f'backbone.{i}.',h,True,update_running=False)
        a.retain_grad(); p.retain_grad(); acts.append((a,p)); h=p
    g=h.mean(dim=2); z=F.linear(g,work['proj.weight'],work['proj.bias']); lo=F.linear(z,work['head.weight'],work['head.bias'])
    loss=O.bce_with_logits(lo,y); loss.backward()
    torch.manual_seed(42)
    m=P.ECGCNN(12,256,5).to(DEV).train()
    store={}
    for i,blk in enumerate(m.backbone):
        def ha(mod,inp,out,i=i):
            out.retain_grad(); store[f'a{i}']=out
        def hp(mod,inp,out,i=i):
            out.retain_grad(); store[f'p{i}']=out
        blk.net[0].register_forward_hook(ha)
        blk.register_forward_hook(hp)
    lg=m(x.to(DEV)); l=Fn.binary_cross_entropy_with_logits(lg,y.to(DEV)); l.backward()
    print(f'B={B} T={T} logits rel {rel(lg,lo):.2e}')
    for i in range(4):
        a,p=acts[i]
        pg = store[f'p{i}'].grad
        print(f'  layer{i}: a {rel(store[f"a{i}"],a):.2e} p {rel(store[f"p{i}"],p):.2e} da {rel(store[f"a{i}"].grad,a.grad):.2e}', 'dp', 'None' if pg is None else f'{rel(pg,p.grad):.2e}')
        # mask mismatch count
        ra = (store[f'a{i}'].grad!=0).cpu(); rb=(a.grad!=0)
    for k,p in m.named_parameters():
        print(f'  {k:32s} {rel(p.grad,leaves[k].grad):.2e}  max|g| {float(leaves[k].grad.abs().max()):.2e}')
