import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200.step import TrainStep
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
kind = sys.argv[4] if len(sys.argv) > 4 else 'cnn'
torch.manual_seed(42)
m = (P.ECGCNN(12,256,5) if kind=='cnn' else P.ECGMultimodal()).cuda().train()
o = P.FusedAdamW(m.parameters(), lr=1.5e-3, weight_decay=1e-4)
e = TrainStep(m, o, B, T, precision=prec)
if os.environ.get('LINEAR'): e.linear = True; e.side = torch.cuda.current_stream()
e.x.normal_(); e.y.bernoulli_(0.3)
if kind!='cnn': e.demo.uniform_()
for _ in range(3): e.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): e.run()
e1.record(); torch.cuda.synchronize()
print('graph step', e0.elapsed_time(e1)*10, 'us')
