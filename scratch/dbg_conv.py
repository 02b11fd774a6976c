import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from ptbxl_multimodal_b200._lib import lib, check, ptr, stream
BF = torch.bfloat16
def to_blocked(x):
    b, c, l = x.shape
    return x.reshape(b, c // 8, 8, l).permute(0, 1, 3, 2).contiguous().to(BF)
def from_blocked(xb, c):
    b, cc, l, _ = xb.shape
    return xb.float().permute(0, 1, 3, 2).reshape(b, c, l)
shapes = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]] or [(40, 16, 32, 1000), (80, 16, 32, 1000), (256, 16, 32, 1000)]
for (B, Ci, Co, L) in shapes:
    torch.manual_seed(0)
    x = torch.randn(B, Ci, L, device='cuda'); w = torch.randn(Co, Ci, 15, device='cuda') * 0.05
    ref = F.conv1d(x.to(BF).float(), w.to(BF).float(), None, padding=7)
    xb = to_blocked(x); wf = torch.empty(15, Ci // 8, Co, 8, dtype=BF, device='cuda')
    check(lib.ecgb200_conv1d_prep_weights_bf16(ptr(w), ptr(wf), None, Co, Ci, stream()), 'prep')
    yb = torch.zeros(B, Co // 8, L, 8, dtype=BF, device='cuda')
    n = lib.ecgb200_conv1d_stat_parts_bf16(B, Ci, Co, L)
    part = torch.zeros(n, 2, Co, device='cuda')
    for _ in range(int(os.environ.get('REPS', '1'))):
        check(lib.ecgb200_conv1d_fwd_stats_bf16(ptr(xb), ptr(wf), None, ptr(yb), ptr(part), B, Ci, Co, L, stream()), 'conv')
    torch.cuda.synchronize()
    y = from_blocked(yb, Co)
    err = float((y - ref).abs().max() / ref.abs().max())
    print(f'B={B} Ci={Ci} Co={Co} L={L} parts={n} rel err {err:.2e}', flush=True)
