"""Latency of the cross-rank kernels, 100 back-to-back launches in a graph (torchrun, >= 2 GPUs)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
from ptbxl_multimodal_b200._lib import lib, check
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 3360 * 300
P_ = symm.empty(n, dtype=torch.float32, device=dev); G_ = symm.empty(n, dtype=torch.float32, device=dev)
F_ = symm.empty(4 * 64, dtype=torch.int32, device=dev); X_ = symm.empty(8 * 512, dtype=torch.float32, device=dev)
P_.normal_(); G_.normal_(); F_.zero_(); X_.zero_()
torch.cuda.synchronize()
hp, hg, hf, hx = (symm.rendezvous(t, dist.group.WORLD) for t in (P_, G_, F_, X_))
W = C.c_void_p * world
ptrs = lambda h, off=0: W(*[int(h.buffer_ptrs[r]) + off for r in range(world)])
M, V = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
hyper = torch.tensor([1.5e-3, 0.9, 0.999, 1e-8, 1e-4, 1.0 / world], device=dev)
step = torch.ones(1, dtype=torch.int32, device=dev)
part = torch.randn(148, 2, 256, device=dev)
out = torch.zeros(world, 2, 256, device=dev)
def timeit(name, fn, iters=100):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(s.cuda_stream); torch.cuda.synchronize(); dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters): fn(s.cuda_stream)
        g.replay(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
    if rank == 0: print(f'{name:50s} {e0.elapsed_time(e1) * 1000 / iters:7.2f} us / launch', flush=True)
    dist.barrier()
timeit('bn_sync C=32 nparts=148', lambda st: check(lib.ecgb200_dp_bn_sync_f32(part.data_ptr(), 148, 32, ptrs(hx), ptrs(hf, 4 * 64 * 2), out.data_ptr(), rank, world, st), 'x'))
timeit('bn_sync C=256 nparts=148', lambda st: check(lib.ecgb200_dp_bn_sync_f32(part.data_ptr(), 148, 256, ptrs(hx), ptrs(hf, 4 * 64 * 2), out.data_ptr(), rank, world, st), 'x'))
timeit('bn_sync C=256 nparts=1', lambda st: check(lib.ecgb200_dp_bn_sync_f32(part.data_ptr(), 1, 256, ptrs(hx), ptrs(hf, 4 * 64 * 2), out.data_ptr(), rank, world, st), 'x'))
for cnt, pad in ((3360, 0), (3360 * 80, 1), (3360 * 150, 3)):
    timeit(f'dp_adamw_fused_range n={cnt}', lambda st: check(lib.ecgb200_dp_adamw_fused_range_f32(ptrs(hp), ptrs(hg), ptrs(hf, 4 * 64 * pad), M.data_ptr(), V.data_ptr(), 0, cnt, rank, world, hyper.data_ptr(), step.data_ptr(), st), 'x'))
I_ = symm.empty(4 * 3360 * 150 + 2 * world * 512, dtype=torch.int64, device=dev); I_.zero_(); torch.cuda.synchronize()
hi = symm.rendezvous(I_, dist.group.WORLD)
ctr = torch.zeros(8, dtype=torch.int32, device=dev)
for cnt in (3360, 3360 * 80, 3360 * 150):
    I_.zero_(); ctr.zero_(); torch.cuda.synchronize(); dist.barrier()
    timeit(f'dp_adamw_ll n={cnt}', lambda st: check(lib.ecgb200_dp_adamw_ll_f32(P_.data_ptr(), G_.data_ptr(), M.data_ptr(), V.data_ptr(), ptrs(hi), ctr.data_ptr(), 0, cnt, rank, world, hyper.data_ptr(), step.data_ptr(), st), 'x'))
I_.zero_(); ctr.zero_(); torch.cuda.synchronize(); dist.barrier()
timeit('bn_sync_ll C=256 nparts=148', lambda st: check(lib.ecgb200_dp_bn_sync_ll_f32(part.data_ptr(), 148, 256, ptrs(hi, 8 * 4 * 3360 * 150), ctr.data_ptr() + 16, out.data_ptr(), rank, world, st), 'x'))
timeit('bn_sync_ll C=32 nparts=148', lambda st: check(lib.ecgb200_dp_bn_sync_ll_f32(part.data_ptr(), 148, 32, ptrs(hi, 8 * 4 * 3360 * 150), ctr.data_ptr() + 16, out.data_ptr(), rank, world, st), 'x'))
x = torch.zeros(1024, device=dev)
timeit('nccl all_reduce 4 KB (for scale)', lambda st: dist.all_reduce(x), iters=20)
dist.barrier(); torch.cuda.synchronize()
import threading, time
threading.Thread(target=lambda: (time.sleep(20), os._exit(0)), daemon=True).start()
del hp, hg, hf, hx, hi
dist.destroy_process_group()
