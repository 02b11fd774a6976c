"""Multi-GPU check of the fused data-parallel step (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py

Same seed and data on both modes; dp_mode='fused' (peer-memory reduce-scatter + AdamW + all-gather kernel)
must give the SAME parameters as dp_mode='nccl' (all-reduce + replicated AdamW): bit-exact for world 2
(a + b is commutative), within 1e-6 relative for larger worlds (different summation order), and all ranks
must hold identical parameters.  Also prints the device time per step of both modes."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ptbxl_multimodal_b200 as P  # noqa: E402
from ptbxl_multimodal_b200.step import TrainStep  # noqa: E402


def run(mode, B, T, steps, rank, dev, kind):
    torch.manual_seed(42)
    model = (P.ECGCNN(12, 256, 5) if kind == "cnn" else P.ECGMultimodal()).to(dev).train()
    opt = P.FusedAdamW(model.parameters(), lr=1.5e-3, weight_decay=1e-4)
    if rank == 0:
        print(f"[{mode} {kind} B={B}] building engine", flush=True)
    eng = TrainStep(model, opt, B, T, precision="bf16", dp_mode=mode, use_graph=os.environ.get("DP_NOGRAPH") is None)
    if rank == 0:
        print(f"[{mode} {kind} B={B}] engine built", flush=True)
    g = torch.Generator().manual_seed(100 + rank)
    xs = [torch.randn(B, 12, T, generator=g).to(dev) for _ in range(steps)]
    ys = [(torch.rand(B, 5, generator=g) < 0.3).float().to(dev) for _ in range(steps)]
    ds = [torch.rand(B, 5, generator=g).to(dev) for _ in range(steps)]
    losses = []
    for i in range(steps):
        losses.append(float(eng(xs[i], ys[i], ds[i] if kind == "mm" else None)))
        if rank == 0:
            print(f"[{mode} {kind} B={B}] step {i} loss {losses[-1]:.5f}", flush=True)
    torch.cuda.synchronize(dev)
    # timing: replay the last batch
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        eng.run()
    e1.record()
    torch.cuda.synchronize(dev)
    return eng, losses, e0.elapsed_time(e1) * 1000 / 50


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for kind, B in (("cnn", 16), ("mm", 8), ("cnn", 128)):
        steps = 3
        ef, lf, tf = run("fused", B, 1000, steps, rank, dev, kind)
        pf = ef.P[:ef.total].clone()          # after steps + 50 replays
        en, ln, tn = run("nccl", B, 1000, steps, rank, dev, kind)
        pn = en.P[:en.total].clone()
        diff = float((pf - pn).abs().max() / pn.abs().max())
        # all ranks identical?
        ref = pf.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, pf))
        lim = 0.0 if world == 2 else 1e-6
        good = diff <= lim and same and all(abs(a - b) <= 1e-6 * max(1.0, abs(b)) for a, b in zip(lf, ln))
        ok = ok and good
        ef.gather_optimizer_state()
        mdiff = float((ef.M[:ef.total] - en.M[:en.total]).abs().max() / en.M[:en.total].abs().max().clamp_min(1e-30))
        ok = ok and mdiff <= max(lim, 1e-6)
        if rank == 0:
            print(f"{kind} B/rank={B} world={world}: fused vs nccl params rel diff {diff:.2e}, moments {mdiff:.2e}, "
                  f"ranks identical {same}, losses {['%.5f' % v for v in lf]} | step fused {tf:.1f} us, nccl {tn:.1f} us "
                  f"-> {'OK' if good else 'MISMATCH'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    good = int(flag) == 1
    if rank == 0:
        print("dp_check passed" if good else "dp_check FAILED", flush=True)
    dist.barrier()
    torch.cuda.synchronize(dev)
    # symmetric-memory mappings + NCCL teardown can block at interpreter exit: leave without running destructors
    sys.stdout.flush()
    os._exit(0 if good else 1)


if __name__ == "__main__":
    main()
