import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200._lib import lib, check, ptr, stream
BF = torch.bfloat16
lib.ecgb200_debug_set_conv_pair(int(os.environ.get('PAIR', '3')))
names = {0: 'start', 1: 'setup done', 2: 'end', 3: 'W resident'}
for g in range(4):
    names[8 + g] = f'x{g} issued'; names[16 + g] = f'x{g} landed'; names[24 + g] = f'acc{g} free'
    names[32 + g] = f'mma{g} issued'; names[40 + g] = f'acc{g} full'; names[48 + g] = f'epi{g} done'
SH = [(256, 64, 128, 250), (256, 128, 256, 125), (256, 256, 128, 125), (256, 128, 64, 250)] if os.environ.get('BIG') else [(256, 16, 32, 1000), (256, 32, 64, 500), (256, 64, 128, 250), (256, 128, 256, 125), (256, 256, 128, 125), (256, 128, 64, 250), (256, 64, 32, 500)]
for (B, Ci, Co, L) in SH:
    xb = torch.randn(B, Ci // 8, L, 8, device='cuda').to(BF)
    wf = (torch.randn(15, Ci // 8, Co, 8, device='cuda') * 0.05).to(BF)
    yb = torch.empty(B, Co // 8, L, 8, dtype=BF, device='cuda')
    part = torch.empty(148, 2, Co, device='cuda')
    tr = torch.zeros(64, dtype=torch.int64, device='cuda')
    for it in range(3):
        check(lib.ecgb200_conv1d_fwd_stats_bf16(ptr(xb), ptr(wf), None, ptr(yb), ptr(part), B, Ci, Co, L, stream()), 'conv')
    torch.cuda.synchronize()
    check(lib.ecgb200_debug_set_trace(tr.data_ptr()), 'trace')
    check(lib.ecgb200_conv1d_fwd_stats_bf16(ptr(xb), ptr(wf), None, ptr(yb), ptr(part), B, Ci, Co, L, stream()), 'conv')
    torch.cuda.synchronize()
    check(lib.ecgb200_debug_set_trace(None), 'trace')
    t = tr.cpu().tolist()
    print(f'--- conv B={B} Ci={Ci} Co={Co} L={L}  (cycles since CTA start; 1965 cyc = 1 us)')
    ev = sorted((v - t[0], names.get(i, str(i))) for i, v in enumerate(t) if v)
    print('  '.join(f'{n}={c}' for c, n in ev))
