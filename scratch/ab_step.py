"""A/B of engine options on the benchmark step: ms/step (median of 5 x 200 replays)."""
import sys, os, statistics, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200.step import TrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
from ptbxl_multimodal_b200._lib import lib
variants = [dict(pair=int(c)) for c in (sys.argv[3] if len(sys.argv) > 3 else '0303')]
for kw in variants:
    kw = dict(kw)
    pm = kw.pop('pair', 3)
    lib.ecgb200_debug_set_conv_pair(pm)
    torch.manual_seed(42)
    m = P.ECGCNN(12, 256, 5).cuda().train()
    o = P.FusedAdamW(m.parameters(), lr=1.5e-3, weight_decay=1e-4)
    e = TrainStep(m, o, B, T, precision='bf16', **kw)
    g = torch.Generator().manual_seed(0)
    for s in (0, 1):
        e.load_batch(torch.randn(B, 12, T, generator=g).cuda(), (torch.rand(B, 5, generator=g) < 0.3).float().cuda(), slot=s)
    for i in range(10): e.run(slot=i & 1)
    torch.cuda.synchronize()
    runs = []
    for r in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200): e.run(slot=i & 1)
        e1.record(); torch.cuda.synchronize()
        runs.append(e0.elapsed_time(e1) / 200)
    print(json.dumps({'opts': kw, 'pair_mask': pm, 'ms_per_step': statistics.median(runs), 'windows_per_s': B / statistics.median(runs) * 1e3,
                      'loss': float(e.loss)}))
