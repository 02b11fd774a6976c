"""Multi-GPU checks of the data-parallel step (run under torchrun, one rank per GPU; tests/test_gpu_dp.py launches it
with 2 ranks when the box has them):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py

1. kernel level  ecgb200_dp_adamw_fused[_range]_f32 == rank-ordered gradient sum + ecgb200_adamw_flat_f32, BIT FOR BIT on every
                 rank, as one bucket and as the engine's two buckets.
2. mode level    dp_mode='fused' (peer-memory kernels, block-4 bucket under backward) == dp_mode='nccl' (all-reduce +
                 replicated AdamW): bit-exact for world 2 (a + b is commutative); all ranks identical.
3. ORACLE, torch-DDP semantics (local BatchNorm): every rank's step against the CPU oracle run on that rank's shard
                 (oracle/ecg_oracle.py: bf16_train_step, the stated bf16 tolerances of tests/test_gpu_step_engine.py), the
                 exchanged update against AdamW on the MEAN of the per-rank oracle gradients, and against
                 torch DistributedDataParallel over the stock modules (oracle/torch_stock.py) on the same GPUs.
4. ORACLE, SyncBN (sync_bn=True): the N-rank step against the single-process oracle on the CONCATENATED batch
                 (SURVEY 8c/8e): logits per shard, mean of the rank gradients == gradient of the whole batch, BatchNorm
                 running statistics of the global batch, loss.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ptbxl_multimodal_b200 as P  # noqa: E402
from ptbxl_multimodal_b200.step import TrainStep  # noqa: E402
from oracle import ecg_oracle as O  # noqa: E402


def rel_inf(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def cos(a, b):
    a = a.detach().double().cpu().flatten(); b = b.detach().double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


def say(rank, *a):
    if rank == 0:
        print(*a, flush=True)


def all_ok(ok, dev):
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return int(flag) == 1


def run(mode, B, T, steps, rank, dev, kind, time_it=True):
    torch.manual_seed(42)
    model = (P.ECGCNN(12, 256, 5) if kind == "cnn" else P.ECGMultimodal()).to(dev).train()
    opt = P.FusedAdamW(model.parameters(), lr=1.5e-3, weight_decay=1e-4)
    eng = TrainStep(model, opt, B, T, precision="bf16", dp_mode=mode)
    g = torch.Generator().manual_seed(100 + rank)
    xs = [torch.randn(B, 12, T, generator=g).to(dev) for _ in range(steps)]
    ys = [(torch.rand(B, 5, generator=g) < 0.3).float().to(dev) for _ in range(steps)]
    ds = [torch.rand(B, 5, generator=g).to(dev) for _ in range(steps)]
    losses = [float(eng(xs[i], ys[i], ds[i] if kind == "mm" else None)) for i in range(steps)]
    torch.cuda.synchronize(dev)
    us = 0.0
    if time_it:
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            eng.run()
        e1.record()
        torch.cuda.synchronize(dev)
        us = e0.elapsed_time(e1) * 1000 / 50
    return eng, losses, us


def kernel_level_check(rank, world, dev):
    import ctypes as C
    import torch.distributed._symmetric_memory as symm
    from ptbxl_multimodal_b200._lib import lib, check
    from ptbxl_multimodal_b200.parallel import padded_size
    n_b, n_a = padded_size(227621), padded_size(491776)
    n = n_b + n_a
    P_ = symm.empty(n, dtype=torch.float32, device=dev)
    G_ = symm.empty(n, dtype=torch.float32, device=dev)
    F_ = symm.empty(3 * 64, dtype=torch.int32, device=dev)
    g0 = torch.Generator().manual_seed(5)
    P_.copy_(torch.randn(n, generator=g0))
    F_.zero_()
    torch.cuda.synchronize(dev)
    hp, hg, hf = (symm.rendezvous(t, dist.group.WORLD) for t in (P_, G_, F_))
    W = C.c_void_p * world
    ptrs = lambda h, off=0: W(*[int(h.buffer_ptrs[r]) + off for r in range(world)])      # noqa: E731
    hyper = torch.tensor([1.5e-3, 0.9, 0.999, 1e-8, 1e-4, 1.0 / world], device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ok = True
    I_ = symm.empty(int(lib.ecgb200_dp_ll_inbox_words(n_a)) + int(lib.ecgb200_dp_ll_inbox_words(n_b)), dtype=torch.int64, device=dev)
    I_.zero_()
    torch.cuda.synchronize(dev)
    hinb = symm.rendezvous(I_, dist.group.WORLD)
    ctr = torch.zeros(4, dtype=torch.int32, device=dev)
    for variant in ("one bucket", "two buckets", "two buckets, one-hop words"):
        P_.copy_(torch.randn(n, generator=torch.Generator().manual_seed(5)))
        M, V = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        Pr, Mr, Vr = P_.clone(), M.clone(), V.clone()
        step = torch.zeros(1, dtype=torch.int32, device=dev)
        for it in range(3):
            gr = torch.Generator().manual_seed(1000 * it + rank)
            G_.copy_(torch.randn(n, generator=gr) * 10.0 ** (-3 * it))
            step += 1
            torch.cuda.synchronize(dev)
            dist.barrier()
            if variant == "one bucket":
                check(lib.ecgb200_dp_adamw_fused_f32(ptrs(hp), ptrs(hg), ptrs(hf), M.data_ptr(), V.data_ptr(), n, rank, world,
                                                     hyper.data_ptr(), step.data_ptr(), st), "dp_adamw_fused")
                owned = [(rank * (n // world), (rank + 1) * (n // world))]
            elif variant == "two buckets, one-hop words":
                owned = []
                boff = 0
                for k, (off, cnt) in enumerate(((n_b, n_a), (0, n_b))):
                    check(lib.ecgb200_dp_adamw_ll_f32(P_.data_ptr(), G_.data_ptr(), M.data_ptr(), V.data_ptr(), ptrs(hinb, boff),
                                                      ctr.data_ptr() + 8 * k, off, cnt, rank, world, hyper.data_ptr(),
                                                      step.data_ptr(), st), "dp_adamw_ll")
                    boff += 8 * int(lib.ecgb200_dp_ll_inbox_words(cnt))
                    owned.append((off + rank * (cnt // world), off + (rank + 1) * (cnt // world)))
            else:
                owned = []
                for off, cnt, pad in ((n_b, n_a, 1), (0, n_b, 2)):
                    check(lib.ecgb200_dp_adamw_fused_range_f32(ptrs(hp), ptrs(hg), ptrs(hf, 4 * 64 * pad), M.data_ptr(), V.data_ptr(),
                                                               off, cnt, rank, world, hyper.data_ptr(), step.data_ptr(), st),
                          "dp_adamw_fused_range")
                    owned.append((off + rank * (cnt // world), off + (rank + 1) * (cnt // world)))
            torch.cuda.synchronize(dev)
            allg = [torch.empty(n, device=dev) for _ in range(world)]
            dist.all_gather(allg, G_.clone())
            gsum = torch.zeros(n, device=dev)
            for r in range(world):
                gsum = gsum + allg[r]
            check(lib.ecgb200_adamw_flat_f32(Pr.data_ptr(), gsum.data_ptr(), Mr.data_ptr(), Vr.data_ptr(), n, hyper.data_ptr(),
                                             step.data_ptr(), st), "adamw_flat")
            torch.cuda.synchronize(dev)
            good = bool(torch.equal(P_, Pr))
            for lo, hi in owned:
                good = good and bool(torch.equal(M[lo:hi], Mr[lo:hi])) and bool(torch.equal(V[lo:hi], Vr[lo:hi]))
            ok = ok and good
        ok = all_ok(ok, dev)
        say(rank, f"1. kernel level ({variant}): fused exchange == rank-ordered sum + AdamW, bit-exact on all {world} ranks: {ok}")
    del hp, hg, hf, hinb
    return ok


def mode_level_check(rank, world, dev):
    ok = True
    for kind, B in (("cnn", 16), ("mm", 8), ("cnn", 128)):
        ef, lf, tf = run("fused", B, 1000, 3, rank, dev, kind)
        pf = ef.P.clone()
        eb, lb, tb = run("barrier", B, 1000, 3, rank, dev, kind)
        same_b = bool(torch.equal(eb.P, pf))                  # one-hop words == flag barriers, bit for bit, any world size
        eb.close()
        del eb
        en, ln, tn = run("nccl", B, 1000, 3, rank, dev, kind)
        pn = en.P.clone()
        diff = float((pf - pn).abs().max() / pn.abs().max())
        ref = pf.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(ref, pf))
        # world 2: a + b is commutative -> the two modes are bit-identical for ever.  world > 2: NCCL sums in a
        # different order, fp32 rounding differs in the last bit and 53 Adam steps through bf16 activations amplify
        # it, so only the first steps' losses are required to agree there (the kernel-level check above is exact).
        lim = 0.0 if world == 2 else float("inf")
        good = diff <= lim and same and same_b and all(abs(a - b) <= 2e-4 * max(1.0, abs(b)) for a, b in zip(lf, ln))
        ef.gather_optimizer_state()
        mdiff = float((ef.M - en.M).abs().max() / en.M.abs().max().clamp_min(1e-30))
        good = good and mdiff <= lim
        ok = ok and good
        say(rank, f"2. {kind} B/rank={B} world={world}: fused vs nccl params rel diff {diff:.2e}, moments {mdiff:.2e}, ranks identical "
                  f"{same}, one-hop == barrier form {same_b}, losses {['%.5f' % v for v in lf]} | step fused {tf:.1f} us, barrier {tb:.1f} us, "
                  f"nccl {tn:.1f} us -> {'OK' if good else 'MISMATCH'}")
        ef.close(); en.close()
        del ef, en
    return all_ok(ok, dev)


def oracle_check(rank, world, dev, kind, sync_bn):
    """One step of the N-rank engine against the CPU oracle (local-BN: per shard + mean of gradients; SyncBN: the
    concatenated batch)."""
    B, T, nl, lr, wd = 8, 1000, 5, 1.5e-3, 1e-4
    shards = [O.synth_batch(B, T, nl, seed=50 + r, with_demo=(kind == "mm")) for r in range(world)]
    x, y = shards[rank][0], shards[rank][-1]
    demo = shards[rank][1] if kind == "mm" else None
    sd = O.init_state_dict(kind, nl, seed=42)
    torch.manual_seed(42)
    model = (P.ECGCNN(12, 256, nl) if kind == "cnn" else P.ECGMultimodal(num_labels=nl)).to(dev).train()
    opt = P.FusedAdamW(model.parameters(), lr=lr, weight_decay=wd)
    eng = TrainStep(model, opt, B, T, precision="bf16", dp_mode="fused", sync_bn=sync_bn, use_graph=False)
    p0 = {k: s.param.detach().clone() for k, s in eng.seg.items()}
    loss = float(eng(x.to(dev), y.to(dev), demo.to(dev) if demo is not None else None))
    torch.cuda.synchronize(dev)
    keys = O.param_keys(sd)
    gl = {k: model.get_parameter(k).grad.detach().clone() for k in keys}       # this rank's LOCAL gradients (fused mode)
    ok = True
    if not sync_bn:
        ref = O.bf16_train_step(sd, x, y, demo=demo)                           # the oracle on this rank's shard
        good = rel_inf(eng.logits, ref["logits"]) < 5e-3 and abs(loss - float(ref["loss"])) < 1e-3 * float(ref["loss"])
        worst = 0.0
        for k in keys:
            if k.endswith("net.0.bias"):
                continue
            c = 1 - cos(gl[k], ref["grads"][k])
            worst = max(worst, c)
            good = good and c < 3e-3
        ok = ok and good
        # the update: AdamW on the mean over ranks of the ORACLE's per-shard gradients (what torch DDP would apply)
        objs = [None] * world
        dist.all_gather_object(objs, {k: v for k, v in ref["grads"].items()})
        gmean = {k: sum(o[k] for o in objs) / world for k in keys}
        sd1 = O.clone_sd(sd)
        O.adamw_step(sd1, gmean, O.AdamWState(sd1, lr, wd))
        # AdamW's first update is -lr * g / (|g| + eps): sign-like, so elements whose gradient is ~0 are ill-conditioned
        # (a 1e-9 difference flips the update by 2 * lr).  The update is therefore compared on the well-conditioned half
        # (|mean oracle gradient| above the tensor's median); the exchanged gradient itself on all elements.
        wu = wg = 0.0
        for k in keys:
            if k.endswith("net.0.bias"):
                continue
            ge = gl[k].clone()
            dist.all_reduce(ge)
            wg = max(wg, 1 - cos(ge / world, gmean[k]))
            du_e = (model.get_parameter(k).detach().cpu() - p0[k].cpu()).flatten()
            du_o = (sd1[k] - sd[k]).flatten()
            big = gmean[k].abs().flatten() >= gmean[k].abs().flatten().median()
            wu = max(wu, 1 - cos(du_e[big], du_o[big]))
        ok = ok and wu < 2e-2 and wg < 3e-3
        say(rank, f"3. oracle, local BN ({kind}): logits rel_inf {rel_inf(eng.logits, ref['logits']):.2e}, worst local-gradient 1-cos "
                  f"{worst:.2e} (<3e-3), mean of rank gradients vs mean of per-shard oracle gradients 1-cos {wg:.2e} (<3e-3), applied "
                  f"update vs AdamW(mean of oracle gradients) on the well-conditioned half 1-cos {wu:.2e} (<2e-2)")
    else:
        xa = torch.cat([s[0] for s in shards]); ya = torch.cat([s[-1] for s in shards])
        da = torch.cat([s[1] for s in shards]) if kind == "mm" else None
        ref = O.bf16_train_step(sd, xa, ya, demo=da)                           # ONE process, the concatenated batch
        ref32 = O.train_step(O.clone_sd(sd), xa, ya, None, demo=da)
        sl = slice(rank * B, (rank + 1) * B)
        good = rel_inf(eng.logits, ref["logits"][sl]) < 5e-3
        lt = torch.tensor([loss], device=dev)
        dist.all_reduce(lt)
        good = good and abs(float(lt) / world - float(ref["loss"])) < 1e-3 * float(ref["loss"])
        worst = 0.0
        for k in keys:
            g = gl[k].clone()
            dist.all_reduce(g)
            g /= world                                                         # mean of the rank gradients
            if k.endswith("net.0.bias"):
                continue
            c = 1 - cos(g, ref["grads"][k])
            worst = max(worst, c)
            good = good and c < 3e-3
        # running statistics of the GLOBAL batch (fp32 oracle on the concatenated batch; bf16 storage tolerance)
        sd32 = O.clone_sd(sd)
        O.train_step(sd32, xa, ya, None, demo=da)
        pre = "ecg_backbone." if kind == "mm" else ""
        wr = 0.0
        for i in range(4):
            for nm in ("running_mean", "running_var"):
                key = f"{pre}backbone.{i}.net.1.{nm}"
                wr = max(wr, rel_inf(model.state_dict()[key], sd32[key]))
        good = good and wr < 2e-2
        ok = ok and good
        say(rank, f"4. oracle, SyncBN ({kind}): logits of shard vs single-process oracle on the concatenated batch rel_inf "
                  f"{rel_inf(eng.logits, ref['logits'][sl]):.2e} (<5e-3), worst 1-cos(mean of rank gradients, whole-batch gradient) "
                  f"{worst:.2e} (<3e-3), running stats rel_inf {wr:.2e} (<2e-2), loss {float(lt) / world:.5f} vs {float(ref['loss']):.5f} "
                  f"(fp32 oracle {float(ref32['loss']):.5f})")
    # all ranks hold the same parameters after the exchange
    ref_p = eng.P.clone()
    dist.broadcast(ref_p, src=0)
    ok = ok and bool(torch.equal(ref_p, eng.P))
    eng.close()
    return all_ok(ok, dev)


def ddp_check(rank, world, dev):
    """The engine (local BN, fp32-exact forward is not available under DP, so bf16) against torch DDP over the stock
    modules in fp32 on the same GPUs: same shards, same init, one step -- update direction and loss."""
    from oracle import torch_stock as S
    B, T, nl, lr, wd = 8, 1000, 5, 1.5e-3, 1e-4
    x, y = O.synth_batch(B, T, nl, seed=50 + rank)
    sd = O.init_state_dict("cnn", nl, seed=42)
    saved = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    stock = S.build("cnn", sd, nl).to(dev).train()
    ddp = torch.nn.parallel.DistributedDataParallel(stock, device_ids=[dev.index])
    opt = torch.optim.AdamW(ddp.parameters(), lr=lr, weight_decay=wd)
    loss_ddp = float(S.make_step(ddp, opt)(x.to(dev), y.to(dev)))
    torch.backends.cudnn.allow_tf32 = saved
    torch.manual_seed(42)
    model = P.ECGCNN(12, 256, nl).to(dev).train()
    eng = TrainStep(model, P.FusedAdamW(model.parameters(), lr=lr, weight_decay=wd), B, T, precision="bf16", use_graph=False)
    loss = float(eng(x.to(dev), y.to(dev)))
    worst = wu = 0.0
    for k, p in stock.named_parameters():
        if k.endswith("net.0.bias"):
            continue
        ge = model.get_parameter(k).grad.detach().clone()
        dist.all_reduce(ge)
        gd = p.grad.detach()                                   # DDP has averaged it over the ranks
        worst = max(worst, 1 - cos(ge / world, gd))
        big = (gd.abs().flatten() >= gd.abs().flatten().median()).cpu()
        du_e = (model.get_parameter(k).detach().cpu() - sd[k]).flatten()
        du_d = (p.detach().cpu() - sd[k]).flatten()
        wu = max(wu, 1 - cos(du_e[big], du_d[big]))
    ok = abs(loss - loss_ddp) < 2e-2 * loss_ddp and worst < 3e-2 and wu < 8e-2
    say(rank, f"3b. torch DDP (stock fp32 modules, NCCL): loss {loss:.5f} vs {loss_ddp:.5f}; mean of rank gradients vs DDP's averaged "
              f"gradient worst 1-cos {worst:.2e} (<3e-2, bf16 vs fp32); applied update on the well-conditioned half 1-cos {wu:.2e} (<8e-2)")
    eng.close()
    return all_ok(ok, dev)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = kernel_level_check(rank, world, dev)
    ok = mode_level_check(rank, world, dev) and ok
    for kind in ("cnn", "mm"):
        ok = oracle_check(rank, world, dev, kind, sync_bn=False) and ok
    ok = ddp_check(rank, world, dev) and ok
    for kind in ("cnn", "mm"):
        ok = oracle_check(rank, world, dev, kind, sync_bn=True) and ok
    good = all_ok(ok, dev)
    say(rank, "dp_check passed" if good else "dp_check FAILED")
    dist.barrier()
    torch.cuda.synchronize(dev)
    sys.stdout.flush()
    # orderly teardown, with a watchdog in case the symmetric-memory / NCCL teardown blocks at exit
    import threading
    import time
    threading.Thread(target=lambda: (time.sleep(30), os._exit(0 if good else 1)), daemon=True).start()
    dist.destroy_process_group()
    sys.exit(0 if good else 1)


if __name__ == "__main__":
    main()
