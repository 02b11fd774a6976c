import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200.step import TrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(42)
m = P.ECGCNN(12, 256, 5).cuda().train()
o = P.FusedAdamW(m.parameters(), lr=1.5e-3, weight_decay=1e-4)
e = TrainStep(m, o, B, 1000, precision='bf16', use_graph=False)
e.x.normal_(); e.y.bernoulli_(0.3)
from ptbxl_multimodal_b200._lib import lib
diag = torch.zeros(64, dtype=torch.int64).pin_memory()
lib.ecgb200_debug_set_diag(diag.data_ptr())
orig = e._k
def k(name, fn, *args):
    try:
        orig(name, fn, *args)
        torch.cuda.synchronize()
    except Exception as ex:
        d = diag.tolist()
        print('diag', hex(d[0]), 'block', d[1] >> 32, 'thread', d[1] & 0xffffffff, 'smem addr', hex(d[2] >> 32), 'parity', d[2] & 0xffffffff)
        for w in range(10):
            if d[4 + 2 * w]:
                print(f'  warp {w}: block {d[4 + 2 * w] >> 32} thread {d[4 + 2 * w] & 0xffffffff} waits smem {hex(d[5 + 2 * w] >> 32)} parity {d[5 + 2 * w] & 0xffffffff}')
        print('FAILED at', name + e._prof_tag, type(ex).__name__, str(ex)[:100], flush=True)
        sys.exit(1)
    print('ok', name + e._prof_tag, flush=True)
e._k = k
e.run()
print('step done')
