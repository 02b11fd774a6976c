import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200.step import TrainStep
for (kind, nl, B, T) in [('cnn', 5, 128, 1000), ('cnn', 5, 64, 1000), ('cnn', 1, 64, 5000), ('mm', 5, 128, 1000), ('cnn', 5, 256, 1000)]:
    torch.manual_seed(42)
    m = (P.ECGCNN(12, 256, nl) if kind == 'cnn' else P.ECGMultimodal()).cuda().train()
    o = P.FusedAdamW(m.parameters(), lr=1.5e-3, weight_decay=1e-4)
    e = TrainStep(m, o, B, T, precision='bf16')
    e.x.normal_(); e.y.bernoulli_(0.3)
    if kind == 'mm': e.demo.uniform_()
    for _ in range(3): e.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): e.run()
    e1.record(); torch.cuda.synchronize()
    print(f'{kind} nl={nl} B={B} T={T}: {e0.elapsed_time(e1) * 10:.1f} us/step, {e.launches_per_step} launches', flush=True)
    del e, m, o
