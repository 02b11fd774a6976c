"""Schedule of the captured bf16 step: which calls overlap on the two streams (TrainStep.trace_schedule)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200.step import TrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
from ptbxl_multimodal_b200._lib import lib
lib.ecgb200_debug_set_conv_pair(int(os.environ.get('PAIR', '3')))
torch.manual_seed(42)
m = P.ECGCNN(12, 256, 5).cuda().train()
o = P.FusedAdamW(m.parameters(), lr=1.5e-3, weight_decay=1e-4)
e = TrainStep(m, o, B, T, precision='bf16')
e.x.normal_(); e.y.bernoulli_(0.3)
for _ in range(3): e.run()
torch.cuda.synchronize()
tl = e.trace_schedule()
for n, sid, a, b in sorted(tl, key=lambda r: r[2]):
    print(f'{"  " * (4 * sid)}[s{sid}] {n:18s} {a:8.1f} -> {b:8.1f}  ({b - a:6.1f} us)')
print('span', max(r[3] for r in tl), 'us')
