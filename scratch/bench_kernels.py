"""Warm, host-overhead-free timing of each C-ABI call of the bf16 step: every call is captured
ITERS times back to back in a CUDA graph and replayed (events around the replay)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200.step import TrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
kind = sys.argv[3] if len(sys.argv) > 3 else 'cnn'
ITERS = 10
from ptbxl_multimodal_b200._lib import lib
lib.ecgb200_debug_set_conv_pair(int(os.environ.get('PAIR', '3')))
torch.manual_seed(42)
m = (P.ECGCNN(12, 256, 5) if kind == 'cnn' else P.ECGMultimodal()).cuda().train()
o = P.FusedAdamW(m.parameters(), lr=1.5e-3, weight_decay=1e-4)
e = TrainStep(m, o, B, T, precision='bf16', use_graph=False)
e.x.normal_(); e.y.bernoulli_(0.3)
if kind != 'cnn': e.demo.uniform_()
for _ in range(3): e.run()
torch.cuda.synchronize()
# record the call list of one step
calls = []
orig = e._k
def rec(name, fn, *args):
    calls.append((name + e._prof_tag, fn, args))
e._k = rec
e._enqueue()
e._k = orig
torch.cuda.synchronize()
main = torch.cuda.current_stream().cuda_stream
tot = 0.0
for name, fn, args in calls:
    args = list(args)
    s = torch.cuda.Stream()
    args[-1] = s.cuda_stream
    with torch.cuda.stream(s):
        for _ in range(2): fn(*args)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(ITERS): fn(*args)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); g.replay(); g.replay(); e1.record(s); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / (2 * ITERS)
    tot += us
    print(f'{name:18s} {us:8.2f} us')
print(f'sum {tot:.1f} us over {len(calls)} calls')
