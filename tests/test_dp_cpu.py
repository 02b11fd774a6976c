"""CPU: the N > 1 path's host logic -- shard plan, C-ABI argument checks, and the reduce-scatter + AdamW +
all-gather protocol of csrc/dp_fused.cu replayed by two gloo processes against all-reduce + full AdamW."""
import ctypes as C
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ptbxl_multimodal_b200 import lib
from ptbxl_multimodal_b200.parallel import DP_MAX_WORLD, padded_size, shard_bounds, split_batch


@pytest.mark.parametrize("n", [1, 31, 32, 719397, 718369, 757221])
def test_shard_plan_covers_the_space(n):
    npad = padded_size(n)
    assert npad >= n and npad % 3360 == 0 and npad - n < 3360
    for world in range(1, DP_MAX_WORLD + 1):
        b = shard_bounds(npad, world)
        assert b[0][0] == 0 and b[-1][1] == npad
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))            # contiguous, no overlap
        assert len({hi - lo for lo, hi in b}) == 1                                # equal
        assert all(lo % 4 == 0 for lo, _ in b)                                    # 16-byte aligned float4 shards
    with pytest.raises(ValueError):
        shard_bounds(npad, DP_MAX_WORLD + 1)
    with pytest.raises(ValueError):
        shard_bounds(npad + 4, 8)                                                 # not a multiple of 4 * world


def test_split_batch():
    assert split_batch(1024, 8) == 128 and split_batch(512, 8) == 64 and split_batch(256, 1) == 256
    with pytest.raises(ValueError):
        split_batch(1000, 3)


def test_dp_kernel_argument_checks_without_gpu():
    assert lib.ecgb200_dp_flag_words(8) == 18
    W = C.c_void_p * 2
    fake = W(16, 32)
    f = lib.ecgb200_dp_adamw_fused_f32
    assert f(None, fake, fake, 16, 16, 64, 0, 2, 16, 16, None) == -1            # NULL table
    assert f(fake, fake, fake, 16, 16, 60, 0, 2, 16, 16, None) == -1            # n % (4 * world) != 0
    assert f(fake, fake, fake, 16, 16, 64, 2, 2, 16, 16, None) == -2            # rank out of range
    W9 = C.c_void_p * 9
    assert f(W9(*[16] * 9), W9(*[16] * 9), W9(*[16] * 9), 16, 16, 72 * 4, 0, 9, 16, 16, None) == -2   # world > 8


def _adamw(p, g, m, v, t, lr=1.5e-3, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4):
    """The arithmetic of adamw_flat_kernel / dp_adamw_fused_kernel (torch.optim.AdamW order of operations)."""
    p = p * (1.0 - lr * wd)
    m = m + (1.0 - b1) * (g - m)
    v = v * b2 + (1.0 - b2) * g * g
    denom = v.sqrt() / (1.0 - b2 ** t) ** 0.5 + eps
    p = p - (lr / (1.0 - b1 ** t)) * (m / denom)
    return p, m, v


def _worker(rank, world, port, n, steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        npad = padded_size(n)
        lo, hi = shard_bounds(npad, world)[rank]
        g0 = torch.Generator().manual_seed(7)
        P = torch.zeros(npad); P[:n] = torch.randn(n, generator=g0)
        Pref, Mref, Vref = P.clone(), torch.zeros(npad), torch.zeros(npad)
        M, V = torch.zeros(npad), torch.zeros(npad)                 # only [lo, hi) is ever touched
        gr = torch.Generator().manual_seed(100 + rank)
        for t in range(1, steps + 1):
            G = torch.zeros(npad); G[:n] = torch.randn(n, generator=gr)
            # ---- protocol of dp_fused.cu: every rank reads every rank's gradient shard in rank order
            allg = [torch.empty(npad) for _ in range(world)]
            dist.all_gather(allg, G)                                 # stands in for NVLink peer loads
            gsum = torch.zeros(hi - lo)
            for r in range(world):
                gsum = gsum + allg[r][lo:hi]
            gsum = gsum * (1.0 / world)
            p_new, M[lo:hi], V[lo:hi] = _adamw(P[lo:hi], gsum, M[lo:hi], V[lo:hi], t)
            shards = [torch.empty(hi - lo) for _ in range(world)]
            dist.all_gather(shards, p_new)                           # stands in for the peer stores
            P = torch.cat(shards)
            # ---- what it must equal: all-reduce(mean) + replicated AdamW
            Gm = G.clone()
            dist.all_reduce(Gm)
            Gm = Gm * (1.0 / world)
            Pref, Mref, Vref = _adamw(Pref, Gm, Mref, Vref, t)
        exact = bool(torch.equal(P, Pref)) if world == 2 else None
        rel = float((P - Pref).abs().max() / Pref.abs().max())
        # sharded moments: gather as TrainStep.gather_optimizer_state does
        for buf, ref in ((M, Mref), (V, Vref)):
            parts = [torch.empty(hi - lo) for _ in range(world)]
            dist.all_gather(parts, buf[lo:hi].clone())
            full = torch.cat(parts)
            rel = max(rel, float((full - ref).abs().max() / ref.abs().max()))
        pad_zero = bool((P[n:] == 0).all())
        q.put((rank, exact, rel, pad_zero))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 719397)])
def test_sharded_optimizer_protocol_equals_allreduce(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, exact, rel, pad_zero in res:
        assert exact is True, f"rank {rank}: world-2 result must be bit-exact (a + b is commutative)"
        assert rel == 0.0 and pad_zero
