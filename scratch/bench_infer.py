"""Eval forward throughput: bf16 engine, fp32x3 (split-precision) engine, fp32 module path (CUDA cores)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
out = {}
for B, T in ((256, 1000), (1024, 1000), (4096, 1000), (512, 5000)):
    torch.manual_seed(42)
    m = P.ECGCNN(12, 256, 5).cuda().eval()
    x = torch.randn(B, 12, T, device='cuda')
    for prec in ('bf16', 'fp32x3'):
        e = P.InferStep(m, B, T, precision=prec)
        e.load_batch(x, slot=0); e.load_batch(x, slot=1); e.capture()
        for i in range(5): e.run(slot=i & 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50 if B * T <= 1024 * 1000 else 20
        e0.record()
        for i in range(n): e.run(slot=i & 1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[f'{prec}_B{B}_T{T}'] = {'ms_per_batch': round(ms, 4), 'windows_per_s': round(B / ms * 1e3)}
        del e
    if B <= 1024:
        with torch.no_grad():
            for _ in range(2): m(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): m(x)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[f'fp32_module_B{B}_T{T}'] = {'ms_per_batch': round(ms, 4), 'windows_per_s': round(B / ms * 1e3)}
print(json.dumps(out, indent=1))
