"""Three un-graphed steps of the benchmarked engine (raw int16 input, bf16 tcgen05 path) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200.step import TrainStep
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
torch.manual_seed(42)
m = P.ECGCNN(12,256,5).cuda().train()
o = P.FusedAdamW(m.parameters(), lr=1.5e-3, weight_decay=1e-4)
e = TrainStep(m, o, 256, 1000, precision=prec, use_graph=False, raw_input=(prec == 'bf16'))
if prec == 'bf16':
    fr = (torch.randn(256, 1000, 12) * 200).round().clamp_(-32767, 32767).to(torch.int16).cuda()
    y = (torch.rand(256, 5) < 0.3).float().cuda()
    e.load_frames(fr, y)
else:
    e.x.normal_(); e.y.bernoulli_(0.3)
for _ in range(3): e.run()
torch.cuda.synchronize()
print('done')
