"""wgrad C-ABI call (tensor kernel + split-K reduce) warm timing per layer at the benchmark shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptbxl_multimodal_b200._lib import lib, check, ptr
BF = torch.bfloat16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for (Ci, Co, L) in [(128, 256, 125), (64, 128, 250), (32, 64, 500), (12, 32, 1000)]:
    Cip = (Ci + 15) // 16 * 16
    xb = torch.randn(B, Cip // 8, L, 8, device='cuda').to(BF)
    dyb = torch.randn(B, Co // 8, L, 8, device='cuda').to(BF)
    dw = torch.empty(Co, Ci, 15, device='cuda'); db = torch.empty(Co, device='cuda')
    ws = torch.empty(lib.ecgb200_conv1d_wgrad_bf16_ws_bytes(B, Ci, Co, L), dtype=torch.uint8, device='cuda')
    s = torch.cuda.Stream()
    def run():
        check(lib.ecgb200_conv1d_wgrad_bf16(ptr(dyb), ptr(xb), ptr(dw), ptr(db), None, 0, ptr(ws), B, Ci, Co, L, s.cuda_stream), 'wgrad')
    with torch.cuda.stream(s):
        for _ in range(3): run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10): run()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); g.replay(); g.replay(); e1.record(s); torch.cuda.synchronize()
    print(f'wgrad Ci={Ci} Co={Co} L={L}: {e0.elapsed_time(e1) * 1000 / 20:.2f} us  (ws {ws.numel() / 1e6:.1f} MB)')
