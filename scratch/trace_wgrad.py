import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptbxl_multimodal_b200._lib import lib, check, ptr, stream
BF = torch.bfloat16
names = {0: 'start', 32: 'all issued', 33: 'acc full', 34: 'epi done'}
for n in range(8):
    names[8 + n] = f'ld{n} issued'; names[16 + n] = f'ld{n} landed'; names[24 + n] = f'mma{n} issued'
for (B, Ci, Co, L) in [(256, 128, 256, 125), (256, 64, 128, 250), (256, 32, 64, 500), (256, 12, 32, 1000)]:
    Cip = (Ci + 15) // 16 * 16
    xb = torch.randn(B, Cip // 8, L, 8, device='cuda').to(BF)
    dyb = torch.randn(B, Co // 8, L, 8, device='cuda').to(BF)
    dw = torch.empty(Co, Ci, 15, device='cuda'); db = torch.empty(Co, device='cuda')
    ws = torch.empty(lib.ecgb200_conv1d_wgrad_bf16_ws_bytes(B, Ci, Co, L), dtype=torch.uint8, device='cuda')
    tr = torch.zeros(64, dtype=torch.int64, device='cuda')
    def run():
        check(lib.ecgb200_conv1d_wgrad_bf16(ptr(dyb), ptr(xb), ptr(dw), ptr(db), None, 0, ptr(ws), B, Ci, Co, L, stream()), 'wgrad')
    for _ in range(3): run()
    torch.cuda.synchronize()
    check(lib.ecgb200_debug_set_trace(tr.data_ptr()), 'trace'); run(); torch.cuda.synchronize()
    check(lib.ecgb200_debug_set_trace(None), 'trace')
    t = tr.cpu().tolist()
    print(f'--- wgrad B={B} Ci={Ci} Co={Co} L={L}')
    ev = sorted((v - t[0], names.get(i, str(i))) for i, v in enumerate(t) if v)
    print('  '.join(f'{n}={c}' for c, n in ev))
