"""Launch ramp / spread / tail of the tcgen05 kernels: per-CTA first/last-instruction %globaltimer stamps, with a
stamp kernel before and after on the same stream."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptbxl_multimodal_b200._lib import lib, check, ptr, stream
BF = torch.bfloat16
def report(name, run, nblk):
    sp = torch.zeros(2 * 4096, dtype=torch.int64, device='cuda')
    st = torch.zeros(4, dtype=torch.int64, device='cuda')
    for _ in range(3): run()
    torch.cuda.synchronize()
    check(lib.ecgb200_debug_set_cta_span(sp.data_ptr()), 'span')
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=s):
        run(); check(lib.ecgb200_debug_stamp(st.data_ptr(), 0, stream()), 's'); run(); check(lib.ecgb200_debug_stamp(st.data_ptr(), 1, stream()), 's')
    g.replay(); g.replay(); torch.cuda.synchronize()
    check(lib.ecgb200_debug_set_cta_span(None), 'span')
    v = sp.cpu().view(-1, 2)[:nblk]
    t0, t1 = st.cpu().tolist()[:2]
    a, b = v[:, 0], v[:, 1]
    live = a > 0
    a, b = a[live], b[live]
    q = lambda t, f: float(torch.quantile(t.double(), f))
    print(f'{name}: stamp-before -> first CTA start {float(a.min() - t0) / 1e3:6.2f} us | CTA starts spread {float(a.max() - a.min()) / 1e3:6.2f} | '
          f'CTA durations min/med/max {float((b - a).min()) / 1e3:6.2f}/{q(b - a, .5) / 1e3:6.2f}/{float((b - a).max()) / 1e3:6.2f} | '
          f'first end {float(b.min() - a.min()) / 1e3:6.2f}  last end {float(b.max() - a.min()) / 1e3:6.2f} | last end -> stamp-after {float(t1 - b.max()) / 1e3:6.2f} | '
          f'total {float(t1 - t0) / 1e3:6.2f} us  ({int(live.sum())} CTAs)')
for (B, Ci, Co, L) in [(256, 16, 32, 1000), (256, 32, 64, 500), (256, 64, 128, 250), (256, 128, 256, 125), (256, 256, 128, 125), (256, 128, 64, 250), (256, 64, 32, 500)]:
    xb = torch.randn(B, Ci // 8, L, 8, device='cuda').to(BF)
    wf = (torch.randn(15, Ci // 8, Co, 8, device='cuda') * 0.05).to(BF)
    yb = torch.empty(B, Co // 8, L, 8, dtype=BF, device='cuda')
    part = torch.empty(148, 2, Co, device='cuda')
    report(f'conv {Ci}->{Co} L={L}', lambda: check(lib.ecgb200_conv1d_fwd_stats_bf16(ptr(xb), ptr(wf), None, ptr(yb), ptr(part), B, Ci, Co, L, stream()), 'conv'), 148)
import os as _os
_os.environ['ECGB200_DEBUG_SKIP_WGRAD_REDUCE'] = '1'
for (B, Ci, Co, L) in [(256, 128, 256, 125), (256, 64, 128, 250), (256, 32, 64, 500), (256, 12, 32, 1000)]:
    Cip = (Ci + 15) // 16 * 16
    xb = torch.randn(B, Cip // 8, L, 8, device='cuda').to(BF)
    dyb = torch.randn(B, Co // 8, L, 8, device='cuda').to(BF)
    dw = torch.empty(Co, Ci, 15, device='cuda'); db = torch.empty(Co, device='cuda')
    ws = torch.empty(lib.ecgb200_conv1d_wgrad_bf16_ws_bytes(B, Ci, Co, L), dtype=torch.uint8, device='cuda')
    report(f'wgrad {Ci}->{Co} L={L} (tc kernel only)', lambda: check(lib.ecgb200_conv1d_wgrad_bf16(ptr(dyb), ptr(xb), ptr(dw), ptr(db), None, 0, ptr(ws), B, Ci, Co, L, stream()), 'wgrad'), 160)
