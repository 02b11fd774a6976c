"""Schedule of the captured data-parallel step (rank 0's view): torchrun --nproc-per-node N scratch/timeline_dp.py [B] [sync_bn]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200.step import TrainStep
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sync = len(sys.argv) > 2 and sys.argv[2] == "sync"
mode = sys.argv[3] if len(sys.argv) > 3 else "fused"
torch.manual_seed(42)
m = P.ECGCNN(12, 256, 5).cuda().train()
o = P.FusedAdamW(m.parameters(), lr=1.5e-3, weight_decay=1e-4)
e = TrainStep(m, o, B, 1000, precision='bf16', sync_bn=sync, dp_mode=mode)
e.x.normal_(); e.y.bernoulli_(0.3)
for _ in range(5): e.run()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(300): e.run()
e1.record(); torch.cuda.synchronize()
if rank == 0: print(f'world {dist.get_world_size()} B/rank {B} sync_bn {sync} dp_mode {mode}: {e0.elapsed_time(e1) * 1000 / 300:.1f} us/step')
dist.barrier()
tl = e.trace_schedule()
if rank == 0:
    for n, sid, a, b in sorted(tl, key=lambda r: r[2]):
        print(f'{"  " * (4 * sid)}[s{sid}] {n:18s} {a:8.1f} -> {b:8.1f}  ({b - a:6.1f} us)')
    print('span', max(r[3] for r in tl), 'us')
dist.barrier(); torch.cuda.synchronize()
e.close()
import threading, time
threading.Thread(target=lambda: (time.sleep(20), os._exit(0)), daemon=True).start()
dist.destroy_process_group()
