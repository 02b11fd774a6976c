"""fp32-input vs raw-int16-input engine: ms/step and schedule (1 GPU)."""
import sys, os, statistics, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200.step import TrainStep
for kind, B, T in (("cnn", 256, 1000), ("mm", 1024, 1000)):
    for raw in (False, True):
        torch.manual_seed(42)
        m = (P.ECGCNN(12, 256, 5) if kind == "cnn" else P.ECGMultimodal()).cuda().train()
        o = P.FusedAdamW(m.parameters(), lr=1.5e-3, weight_decay=1e-4)
        e = TrainStep(m, o, B, T, precision='bf16', raw_input=raw)
        g = torch.Generator().manual_seed(0)
        for s in (0, 1):
            x = torch.randn(B, 12, T, generator=g)
            y = (torch.rand(B, 5, generator=g) < 0.3).float().cuda()
            d = torch.rand(B, 5, generator=g).cuda() if kind == "mm" else None
            if raw:
                e.load_frames((x.transpose(1, 2) * 200).round().clamp_(-32767, 32767).to(torch.int16).contiguous().cuda(), y, d, slot=s)
            else:
                e.load_batch(x.cuda(), y, d, slot=s)
        for i in range(10): e.run(slot=i & 1)
        torch.cuda.synchronize()
        runs = []
        for r in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(100): e.run(slot=i & 1)
            e1.record(); torch.cuda.synchronize()
            runs.append(e0.elapsed_time(e1) / 100)
        print(json.dumps({'kind': kind, 'B': B, 'raw': raw, 'ms_per_step': round(statistics.median(runs), 4), 'runs': [round(r, 4) for r in runs]}), flush=True)
        if B == 1024:
            tl = e.trace_schedule()
            for n, sid, a, b in sorted(tl, key=lambda r: r[2]):
                print(f'{"  " * (4 * sid)}[s{sid}] {n:18s} {a:8.1f} -> {b:8.1f}  ({b - a:6.1f} us)')
        del e, m, o
