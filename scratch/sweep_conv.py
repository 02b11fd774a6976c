import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptbxl_multimodal_b200._lib import lib, check, ptr, stream
BF=torch.bfloat16; DEV='cuda:0'
B=256
cfgs=[('fwdL1',16,32,1000),('fwdL2',32,64,500),('fwdL3',64,128,250),('fwdL4',128,256,125),('dgrL4',256,128,125),('dgrL3',128,64,250),('dgrL2',64,32,500)]
R=os.environ.get('ECGB200_CONV_R','default')
out=[]
for name,Ci,Co,L in cfgs:
    xb=torch.randn(B,Ci//8,L,8,device=DEV).to(BF); wf=torch.randn(15,Ci//8,Co,8,device=DEV).to(BF)*0.05
    yb=torch.empty(B,Co//8,L,8,dtype=BF,device=DEV); bias=torch.zeros(Co,device=DEV)
    flush=torch.empty(256*1024*1024//4,device=DEV)
    ts=[]
    for it in range(6):
        flush.zero_()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); check(lib.ecgb200_conv1d_fwd_bf16(ptr(xb),ptr(wf),ptr(bias),ptr(yb),B,Ci,Co,L,stream()),'c'); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1000)
    t=sorted(ts[1:])[len(ts[1:])//2]
    fl=2*B*L*Co*Ci*15
    out.append(f'{name}:{t:6.1f}us({fl/t/1e6:5.0f}TF)')
print('R='+R, ' '.join(out))
