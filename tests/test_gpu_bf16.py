"""GPU: bf16 tensor-core (tcgen05/TMEM/TMA) kernels through the C ABI against fp32 math on
bf16-rounded operands.  Stated tolerance: operands are rounded to bf16 (2^-9 relative),
products accumulate in fp32, outputs are rounded to bf16 once -> rel_inf <= 1e-2 per tensor
against the fp32 result computed from the same rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from ptbxl_multimodal_b200._lib import lib, check, ptr, stream

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF = torch.bfloat16


def rel_inf(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def gen(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def to_blocked(x):                      # (B,C,L) fp32 -> [B][C/8][L][8] bf16 (reference re-layout in torch)
    b, c, l = x.shape
    return x.reshape(b, c // 8, 8, l).permute(0, 1, 3, 2).contiguous().to(BF)


def from_blocked(xb, c):                # inverse
    b, cc, l, _ = xb.shape
    return xb.float().permute(0, 1, 3, 2).reshape(b, c, l)


def test_pack_unpack_roundtrip():
    x = gen(3, 12, 250, seed=1)
    xg = x.to(DEV)
    xb = torch.empty(3, 2, 250, 8, dtype=BF, device=DEV)
    check(lib.ecgb200_pack_input_bf16(ptr(xg), ptr(xb), 3, 12, 250, stream()), "pack")
    ref = torch.zeros(3, 16, 250); ref[:, :12] = x
    assert torch.equal(xb.cpu(), to_blocked(ref))
    out = torch.empty(3, 16, 250, device=DEV)
    check(lib.ecgb200_unpack_act_bf16(ptr(xb), ptr(out), 3, 16, 250, stream()), "unpack")
    assert torch.equal(out.cpu(), ref.to(BF).float())


@pytest.mark.parametrize("B,Ci,Co,L", [(2, 16, 32, 1000), (2, 32, 64, 500), (3, 64, 128, 250), (2, 128, 256, 125),
                                       (1, 128, 256, 625), (1, 16, 32, 40), (2, 64, 128, 129)])
def test_conv1d_fwd_bf16(B, Ci, Co, L):
    x = gen(B, Ci, L, seed=2)
    w = gen(Co, Ci, 15, seed=3, scale=0.05)
    bias = gen(Co, seed=4, scale=0.1)
    xr, wr = x.to(BF).float(), w.to(BF).float()
    ref = F.conv1d(xr, wr, bias, padding=7)
    xb = to_blocked(x).to(DEV)
    wg = w.to(DEV)
    wf = torch.empty(15, Ci // 8, Co, 8, dtype=BF, device=DEV)
    wd = torch.empty(15, Co // 8, Ci, 8, dtype=BF, device=DEV)
    check(lib.ecgb200_conv1d_prep_weights_bf16(ptr(wg), ptr(wf), ptr(wd), Co, Ci, stream()), "prep")
    assert torch.equal(wf.cpu(), wr.permute(2, 1, 0).reshape(15, Ci // 8, 8, Co).permute(0, 1, 3, 2).contiguous().to(BF))
    yb = torch.full((B, Co // 8, L, 8), float("nan"), dtype=BF, device=DEV)
    bias_g = bias.to(DEV)
    check(lib.ecgb200_conv1d_fwd_bf16(ptr(xb), ptr(wf), ptr(bias_g), ptr(yb), B, Ci, Co, L, stream()), "conv_tc")
    torch.cuda.synchronize()
    y = from_blocked(yb.cpu(), Co)
    assert torch.isfinite(y).all()
    assert rel_inf(y, ref) < 1e-2, rel_inf(y, ref)
    if Ci % 32:
        return                      # layer 1 (12->16 padded leads) needs no input gradient
    # dgrad = same kernel with the flipped/transposed weights and no bias
    dy = gen(B, Co, L, seed=5)
    dyr = dy.to(BF).float()
    xr_ = xr.clone().requires_grad_(True)
    F.conv1d(xr_, wr, None, padding=7).backward(dyr)
    dxb = torch.full((B, Ci // 8, L, 8), float("nan"), dtype=BF, device=DEV)
    dyb = to_blocked(dy).to(DEV)
    check(lib.ecgb200_conv1d_fwd_bf16(ptr(dyb), ptr(wd), None, ptr(dxb), B, Co, Ci, L, stream()), "dgrad_tc")
    torch.cuda.synchronize()
    assert rel_inf(from_blocked(dxb.cpu(), Ci), xr_.grad) < 1e-2


@pytest.mark.parametrize("B,Ci,Co,L", [(4, 12, 32, 1000), (3, 32, 64, 500), (5, 64, 128, 250), (6, 128, 256, 125),
                                       (2, 128, 256, 625), (1, 32, 64, 40)])
def test_conv1d_wgrad_bf16(B, Ci, Co, L):
    Cip = (Ci + 15) // 16 * 16
    x = torch.zeros(B, Cip, L); x[:, :Ci] = gen(B, Ci, L, seed=6)
    dy = gen(B, Co, L, seed=7)
    xr = x.to(BF).float()[:, :Ci]
    dyr = dy.to(BF).float()
    w = torch.zeros(Co, Ci, 15, requires_grad=True)
    F.conv1d(xr, w, None, padding=7).backward(dyr)
    ws = torch.empty(lib.ecgb200_conv1d_wgrad_bf16_ws_bytes(B, Ci, Co, L), dtype=torch.uint8, device=DEV)
    dw = torch.full((Co, Ci, 15), float("nan"), device=DEV)
    db = torch.full((Co,), float("nan"), device=DEV)
    dbp = dyr.sum(dim=2).t().contiguous().to(DEV)                  # [Co][B] per-sample sums of dy
    dyb, xb = to_blocked(dy).to(DEV), to_blocked(x).to(DEV)         # keep alive: the call is asynchronous
    check(lib.ecgb200_conv1d_wgrad_bf16(ptr(dyb), ptr(xb), ptr(dw), ptr(db),
                                        ptr(dbp), B, ptr(ws), B, Ci, Co, L, stream()), "wgrad_tc")
    torch.cuda.synchronize()
    assert rel_inf(dw, w.grad) < 2e-3, rel_inf(dw, w.grad)       # fp32 accumulation of exact bf16 products
    assert rel_inf(db, dyr.sum(dim=(0, 2))) < 1e-4


def _bn_ref(y, gamma, beta, train, use_gap, dout_seed=10):
    C = y.shape[1]
    rm, rv = torch.zeros(C), torch.ones(C)
    h = F.batch_norm(y, rm, rv, gamma, beta, training=train, momentum=0.1, eps=1e-5)
    p = F.max_pool1d(F.relu(h), 2)
    out = p.mean(dim=2) if use_gap else p
    dout = gen(*out.shape, seed=dout_seed)
    if not use_gap:
        dout = dout.to(BF).float()
    out.backward(dout)
    return p, out, dout, rm, rv


@pytest.mark.parametrize("B,C,L", [(4, 32, 1000), (3, 64, 250), (5, 256, 125), (2, 128, 31)])
@pytest.mark.parametrize("use_gap", [False, True])
def test_bn_relu_pool_bf16(B, C, L, use_gap):
    y = (gen(B, C, L, seed=5) * 1.7 + 0.3).to(BF).float().requires_grad_(True)      # exactly representable
    gamma = (1 + 0.2 * gen(C, seed=6)).requires_grad_(True)
    beta = (0.1 * gen(C, seed=7)).requires_grad_(True)
    p_ref, out_ref, dout, rm, rv = _bn_ref(y, gamma, beta, True, use_gap)
    yb = to_blocked(y.detach()).to(DEV)
    gg, bg = gamma.detach().to(DEV), beta.detach().to(DEV)
    rmg, rvg = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    st = torch.empty(4, C, device=DEV)
    ws = torch.empty(lib.ecgb200_bn_bwd_ws_bytes(B, C), dtype=torch.uint8, device=DEV)
    check(lib.ecgb200_bn_train_stats_bf16(ptr(yb), ptr(gg), ptr(bg), ptr(rmg), ptr(rvg), ptr(nbt), ptr(st), ptr(ws),
                                          B, C, L, 0.1, 1e-5, stream()), "bn_stats_bf16")
    assert rel_inf(rmg, rm) < 1e-5 and rel_inf(rvg, rv) < 1e-5 and int(nbt) == 1
    Lp = L // 2
    pb = torch.empty(B, C // 8, Lp, 8, dtype=BF, device=DEV)
    gap = torch.empty(B, C, device=DEV) if use_gap else None
    check(lib.ecgb200_bn_relu_pool_fwd_bf16(ptr(yb), ptr(st), ptr(pb), ptr(gap), B, C, L, stream()), "bn_fwd_bf16")
    assert rel_inf(from_blocked(pb.cpu(), C), p_ref) < 6e-3          # one bf16 rounding of the output
    if use_gap:
        assert rel_inf(gap, out_ref) < 1e-5                          # gap accumulates the unrounded fp32 values
    dyb = torch.empty(B, C // 8, L, 8, dtype=BF, device=DEV)
    dgm, dbt = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    dbp = torch.empty(C, lib.ecgb200_bn_nsplit(B, C), device=DEV)
    dpb = None if use_gap else to_blocked(dout).to(DEV)
    dgap = dout.to(DEV) if use_gap else None
    check(lib.ecgb200_bn_relu_pool_bwd_bf16(ptr(yb), ptr(st), ptr(dpb), ptr(dgap), ptr(dyb), ptr(dgm), ptr(dbt),
                                            ptr(dbp), ptr(ws), B, C, L, 1, stream()), "bn_bwd_bf16")
    torch.cuda.synchronize()
    assert rel_inf(dgm, gamma.grad) < 1e-4 and rel_inf(dbt, beta.grad) < 1e-4
    dy = from_blocked(dyb.cpu(), C)
    assert rel_inf(dy, y.grad) < 6e-3
    assert float((dbp.sum(dim=1).cpu() - dy.sum(dim=(0, 2))).abs().max()) < 1e-2 * float(dy.abs().max()) * (B * L) ** 0.5


# --------------------------------------------------------------------------- fused per-step kernels
@pytest.mark.parametrize("B,Ci,Co,L", [(8, 16, 32, 1000), (5, 32, 64, 500), (300, 32, 64, 250), (7, 64, 128, 250),
                                       (3, 128, 256, 125), (200, 128, 256, 125), (2, 64, 128, 129)])
def test_conv_stats_and_fused_bn_forward(B, Ci, Co, L):
    """conv epilogue statistics + one-pass BN-finalise/ReLU/pool against torch on the bf16-rounded conv
    output; persistent multi-group schedule (B large) and partial groups / ragged tiles (B, L odd)."""
    x = gen(B, Ci, L, seed=12)
    w = gen(Co, Ci, 15, seed=13, scale=0.05)
    bias = gen(Co, seed=14, scale=0.1)
    gamma = 1 + 0.2 * gen(Co, seed=15)
    beta = 0.1 * gen(Co, seed=16)
    xr, wr = x.to(BF).float(), w.to(BF).float()
    yref = F.conv1d(xr, wr, bias, padding=7).to(BF).float()            # what the kernels store
    rm, rv = torch.zeros(Co), torch.ones(Co)
    h = F.batch_norm(yref, rm, rv, gamma, beta, training=True, momentum=0.1, eps=1e-5)
    pref = F.max_pool1d(F.relu(h), 2)
    xb, wg, bg = to_blocked(x).to(DEV), w.to(DEV), bias.to(DEV)
    wf = torch.empty(15, Ci // 8, Co, 8, dtype=BF, device=DEV)
    check(lib.ecgb200_conv1d_prep_weights_bf16(ptr(wg), ptr(wf), None, Co, Ci, stream()), "prep")
    nparts = lib.ecgb200_conv1d_stat_parts_bf16(B, Ci, Co, L)
    assert 0 < nparts <= torch.cuda.get_device_properties(0).multi_processor_count
    part = torch.full((nparts, 2, Co), float("nan"), device=DEV)
    yb = torch.full((B, Co // 8, L, 8), float("nan"), dtype=BF, device=DEV)
    check(lib.ecgb200_conv1d_fwd_stats_bf16(ptr(xb), ptr(wf), ptr(bg), ptr(yb), ptr(part), B, Ci, Co, L, stream()), "conv")
    torch.cuda.synchronize()
    y = from_blocked(yb.cpu(), Co)
    assert torch.isfinite(y).all() and torch.isfinite(part).all()
    assert rel_inf(y, yref) < 1e-2
    s1, s2 = part[:, 0].double().sum(0).cpu(), part[:, 1].double().sum(0).cpu()
    assert rel_inf(s1, y.double().sum(dim=(0, 2))) < 1e-4 * (B * L) ** 0.5      # stats of the stored values
    assert rel_inf(s2, (y.double() ** 2).sum(dim=(0, 2))) < 1e-5
    gg, btg = gamma.to(DEV), beta.to(DEV)
    rmg, rvg = torch.zeros(Co, device=DEV), torch.ones(Co, device=DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    st = torch.empty(4, Co, device=DEV)
    for use_gap in (False, True):
        pb = torch.full((B, Co // 8, L // 2, 8), float("nan"), dtype=BF, device=DEV)
        gap = torch.empty(B, Co, device=DEV) if use_gap else None
        check(lib.ecgb200_bn_relu_pool_fwd_train_bf16(ptr(yb), ptr(part), nparts, ptr(gg), ptr(btg), ptr(rmg), ptr(rvg),
                                                      ptr(nbt), ptr(st), ptr(pb), ptr(gap), B, Co, L, 0.1, 1e-5,
                                                      1, stream()), "bn_fwd_train")
        torch.cuda.synchronize()
        # reference statistics from the values the GPU actually stored (y differs from yref by bf16 ties)
        rm2, rv2 = torch.zeros(Co), torch.ones(Co)
        h2 = F.batch_norm(y, rm2, rv2, gamma, beta, training=True, momentum=0.1, eps=1e-5)
        p2 = F.max_pool1d(F.relu(h2), 2)
        assert rel_inf(from_blocked(pb.cpu(), Co), p2) < 6e-3
        if use_gap:
            assert rel_inf(gap, p2.mean(dim=2)) < 2e-5
        else:
            assert rel_inf(rmg, rm2) < 1e-5 and rel_inf(rvg, rv2) < 1e-4 and int(nbt) == 1
            mean = y.mean(dim=(0, 2)); var = y.var(dim=(0, 2), unbiased=False)
            assert rel_inf(st[0], mean) < 1e-5 and rel_inf(st[1], 1 / torch.sqrt(var + 1e-5)) < 1e-4
    assert rel_inf(pref, p2) < 2e-2


@pytest.mark.parametrize("B,NL", [(256, 5), (7, 1), (33, 5)])
def test_fused_head(B, NL):
    """head_fwd_bwd + head_wgrad == proj -> head -> BCE and its autograd (ecg_cnn.py:63-64, loop.py:32-33)."""
    Cin = F_ = 256
    g = torch.Generator().manual_seed(3)
    gap = torch.rand(B, Cin, generator=g)
    wp = (torch.rand(F_, Cin, generator=g) - 0.5) / 8
    bp = (torch.rand(F_, generator=g) - 0.5) / 8
    wh = (torch.rand(NL, F_, generator=g) - 0.5) / 8
    bh = (torch.rand(NL, generator=g) - 0.5) / 8
    y = (torch.rand(B, NL, generator=g) < 0.3).float()
    leaves = [t.clone().requires_grad_(True) for t in (gap, wp, bp, wh, bh)]
    z = F.linear(leaves[0], leaves[1], leaves[2])
    z.retain_grad()
    logits = F.linear(z, leaves[3], leaves[4])
    loss = F.binary_cross_entropy_with_logits(logits, y)
    loss.backward()
    d = lambda t: t.to(DEV).contiguous()          # noqa: E731
    gap_g, wp_g, bp_g, wh_g, bh_g, y_g = map(d, (gap, wp, bp, wh, bh, y))
    wpT = wp_g.t().contiguous()
    e = lambda *s: torch.full(s, float("nan"), device=DEV)       # noqa: E731
    zz, lg, dl, dz, dgap = e(B, F_), e(B, NL), e(B, NL), e(B, F_), e(B, Cin)
    lp = e(lib.ecgb200_head_loss_parts(B))
    check(lib.ecgb200_head_fwd_bwd_f32(ptr(gap_g), ptr(wpT), ptr(wp_g), ptr(bp_g), ptr(wh_g), ptr(bh_g), ptr(y_g),
                                       ptr(zz), ptr(lg), ptr(dl), ptr(dz), ptr(dgap), ptr(lp), B, Cin, F_, NL, 1.0,
                                       stream()), "head_fwd_bwd")
    dwp, dbp, dwh, dbh, ls = e(F_, Cin), e(F_), e(NL, F_), e(NL), e(1)
    check(lib.ecgb200_head_wgrad_f32(ptr(gap_g), ptr(zz), ptr(dz), ptr(dl), ptr(lp), ptr(dwp), ptr(dbp), ptr(dwh),
                                     ptr(dbh), ptr(ls), B, Cin, F_, NL, stream()), "head_wgrad")
    torch.cuda.synchronize()
    tol = 1e-5
    assert rel_inf(lg, logits) < tol and rel_inf(zz, z) < tol
    assert abs(float(ls) - float(loss)) < 1e-6 * max(1.0, abs(float(loss)))
    assert rel_inf(dz, z.grad) < tol and rel_inf(dgap, leaves[0].grad) < tol
    assert rel_inf(dwp, leaves[1].grad) < tol and rel_inf(dbp, leaves[2].grad) < tol
    assert rel_inf(dwh, leaves[3].grad) < tol and rel_inf(dbh, leaves[4].grad) < tol


def test_step_prep_and_flat_adamw():
    import ctypes as C
    B, T = 3, 250
    chan = [12, 32, 64, 128, 256]
    x = gen(B, 12, T, seed=21).to(DEV)
    ws = [gen(chan[l + 1], chan[l], 15, seed=30 + l, scale=0.1).to(DEV) for l in range(4)]
    cip = [16, 32, 64, 128]
    wf = [torch.empty(15, cip[l] // 8, chan[l + 1], 8, dtype=BF, device=DEV) for l in range(4)]
    wd = [None] + [torch.empty(15, chan[l + 1] // 8, cip[l], 8, dtype=BF, device=DEV) for l in range(1, 4)]
    wp = gen(256, 256, seed=40).to(DEV)
    wpT = torch.empty(256, 256, device=DEV)
    ctr = torch.tensor([4], dtype=torch.int32, device=DEV)
    xb = torch.empty(B, 2, T, 8, dtype=BF, device=DEV)
    PV, I4 = C.c_void_p * 4, C.c_int * 4
    check(lib.ecgb200_step_prep_bf16(ptr(x), ptr(xb), B, 12, T, 4, PV(*[ptr(w) for w in ws]), PV(*[ptr(w) for w in wf]),
                                     PV(*[ptr(w) for w in wd]), I4(*chan[1:]), I4(*chan[:4]), ptr(wp), ptr(wpT), 256, 256,
                                     ptr(ctr), stream()), "step_prep")
    xb2 = torch.empty_like(xb)
    check(lib.ecgb200_pack_input_bf16(ptr(x), ptr(xb2), B, 12, T, stream()), "pack")
    assert torch.equal(xb, xb2) and int(ctr) == 5 and torch.equal(wpT, wp.t())
    for l in range(4):
        f2 = torch.empty_like(wf[l])
        d2 = torch.empty_like(wd[l]) if wd[l] is not None else None
        check(lib.ecgb200_conv1d_prep_weights_bf16(ptr(ws[l]), ptr(f2), ptr(d2), chan[l + 1], chan[l], stream()), "prep")
        assert torch.equal(wf[l], f2) and (d2 is None or torch.equal(wd[l], d2))
    # flat AdamW == multi-tensor AdamW at the same step index
    n = 10007 * 4
    p0, g0 = gen(n, seed=50).to(DEV), gen(n, seed=51).to(DEV)
    hyper = torch.tensor([1.5e-3, 0.9, 0.999, 1e-8, 1e-4, 0.5], device=DEV)
    pa, ma, va = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    pb_, mb, vb = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    ca = torch.tensor([0], dtype=torch.int32, device=DEV)
    cb = torch.tensor([1], dtype=torch.int32, device=DEV)
    one = C.c_void_p * 1
    for it in range(3):
        check(lib.ecgb200_adamw_f32(1, one(ptr(pa)), one(ptr(g0)), one(ptr(ma)), one(ptr(va)), (C.c_int64 * 1)(n),
                                    ptr(hyper), ptr(ca), stream()), "adamw")
        check(lib.ecgb200_adamw_flat_f32(ptr(pb_), ptr(g0), ptr(mb), ptr(vb), n, ptr(hyper), ptr(cb), stream()), "adamw_flat")
        cb += 1
    torch.cuda.synchronize()
    assert torch.equal(pa, pb_) and torch.equal(ma, mb) and torch.equal(va, vb)


@pytest.mark.parametrize("B,C,L", [(4, 32, 1000), (6, 64, 250), (4, 256, 125), (2, 128, 31), (256, 64, 500), (64, 256, 125)])
@pytest.mark.parametrize("use_gap", [False, True])
def test_bn_backward_split_passes_and_syncbn_on_one_gpu(B, C, L, use_gap):
    """(a) The reduce and apply passes as separate C-ABI calls == the combined call, bit for bit.
    (b) SyncBN arithmetic emulated on one GPU: the batch is cut into two "replicas"; each runs pass 1 on its half,
    the per-replica pairs are gathered in rank order (what ecgb200_dp_bn_sync_f32 does over peer memory) and pass 2 runs
    per half with nrep = 2 -- the halves' dy equal the single-device result on the whole batch, and the per-replica
    dgamma / dbeta add up to the whole batch's (the gradient exchange sums them)."""
    y = (gen(B, C, L, seed=5) * 1.7 + 0.3)
    gamma, beta = (1 + 0.2 * gen(C, seed=6)).to(DEV), (0.1 * gen(C, seed=7)).to(DEV)
    yb = to_blocked(y).to(DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    st = torch.empty(4, C, device=DEV)
    ws = torch.empty(lib.ecgb200_bn_bwd_ws_bytes(B, C), dtype=torch.uint8, device=DEV)
    check(lib.ecgb200_bn_train_stats_bf16(ptr(yb), ptr(gamma), ptr(beta), ptr(rm), ptr(rv), ptr(nbt), ptr(st), ptr(ws),
                                          B, C, L, 0.1, 1e-5, stream()), "stats")
    Lp = L // 2
    dout = gen(B, C, seed=11) if use_gap else gen(B, C, Lp, seed=11)
    dpb = None if use_gap else to_blocked(dout).to(DEV)
    dgap = dout.to(DEV) if use_gap else None

    def combined(yb_, dpb_, dgap_, b):
        dyb = torch.full((b, C // 8, L, 8), float("nan"), dtype=BF, device=DEV)
        dgm, dbt = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
        dbp = torch.empty(C, lib.ecgb200_bn_nsplit(b, C), device=DEV)
        check(lib.ecgb200_bn_relu_pool_bwd_bf16(ptr(yb_), ptr(st), ptr(dpb_), ptr(dgap_), ptr(dyb), ptr(dgm), ptr(dbt), ptr(dbp),
                                                ptr(ws), b, C, L, 1, stream()), "bn_bwd")
        torch.cuda.synchronize()
        return dyb, dgm, dbt, dbp

    def reduce(yb_, dpb_, dgap_, b):
        ns = lib.ecgb200_bn_nsplit(b, C)
        part = torch.full((ns, 2, C), float("nan"), device=DEV)
        check(lib.ecgb200_bn_relu_pool_bwd_reduce_bf16(ptr(yb_), ptr(st), ptr(dpb_), ptr(dgap_), ptr(part), b, C, L, stream()),
              "bn_bwd_reduce")
        return part

    def apply(yb_, dpb_, dgap_, b, part, local_idx, nrep):
        dyb = torch.full((b, C // 8, L, 8), float("nan"), dtype=BF, device=DEV)
        dgm, dbt = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
        dbp = torch.empty(C, lib.ecgb200_bn_nsplit(b, C), device=DEV)
        check(lib.ecgb200_bn_relu_pool_bwd_apply_bf16(ptr(yb_), ptr(st), ptr(dpb_), ptr(dgap_), ptr(part), part.shape[0],
                                                      local_idx, nrep, ptr(dyb), ptr(dgm), ptr(dbt), ptr(dbp), b, C, L, 1,
                                                      stream()), "bn_bwd_apply")
        torch.cuda.synchronize()
        return dyb, dgm, dbt, dbp

    dy0, g0, b0, s0 = combined(yb, dpb, dgap, B)
    # (a) split passes
    part = reduce(yb, dpb, dgap, B)
    dy1, g1, b1, s1 = apply(yb, dpb, dgap, B, part, -1, 1)
    assert torch.equal(dy0, dy1) and torch.equal(g0, g1) and torch.equal(b0, b1) and torch.equal(s0, s1)
    # (b) two replicas of B/2 windows each, statistics (bn_state) of the whole batch
    h = B // 2
    halves = []
    for r in range(2):
        sl = slice(r * h, (r + 1) * h)
        halves.append((yb[sl].contiguous(), None if use_gap else dpb[sl].contiguous(), dgap[sl].contiguous() if use_gap else None))
    pairs = torch.stack([reduce(*hv, h).double().sum(dim=0).float() for hv in halves])      # (2, 2, C): one pair per replica
    # the whole batch's upstream gradient is what each replica sees (same loss scaling): compare directly
    gsum, bsum = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    for r, hv in enumerate(halves):
        dyr, gr, br, _ = apply(*hv, h, pairs, r, 2)
        ref = dy0[r * h:(r + 1) * h].float()
        assert torch.isfinite(dyr.float()).all()
        assert rel_inf(dyr.float().cpu(), ref.cpu()) < 8e-3                 # same math, another summation order: bf16 ties
        assert float((dyr.float() != ref).float().mean()) < 2e-3
        gsum += gr; bsum += br
    assert rel_inf(gsum, g0) < 2e-5 and rel_inf(bsum, b0) < 2e-5


@pytest.mark.parametrize("B", [256, 7])
def test_fused_multimodal_head(B):
    """mm_head_fwd_bwd + head_wgrad_multi == the FiLM head of ECGMultimodal.forward (ecg_multimodal.py:88-99) + BCE and
    its autograd, for every parameter and for gap."""
    Cin = F_ = 256; D, H, NL = 5, 64, 5
    g = torch.Generator().manual_seed(4)
    r = lambda *s: (torch.rand(*s, generator=g) - 0.5)        # noqa: E731
    gap, demo = torch.rand(B, Cin, generator=g), torch.rand(B, D, generator=g)
    prm = dict(wp=r(F_, Cin) / 8, bp=r(F_) / 8, w0=r(H, D), b0=r(H), w2=r(H, H) / 4, b2=r(H) / 4, wf=r(2 * F_, H) / 4,
               bf=r(2 * F_) / 4, wh=r(NL, F_) / 8, bh=r(NL) / 8)
    y = (torch.rand(B, NL, generator=g) < 0.3).float()
    L = {k: v.clone().requires_grad_(True) for k, v in prm.items()}
    gl = gap.clone().requires_grad_(True)
    h1 = torch.relu(F.linear(demo, L["w0"], L["b0"]))
    h2 = torch.relu(F.linear(h1, L["w2"], L["b2"]))
    film = F.linear(h2, L["wf"], L["bf"])
    z = F.linear(gl, L["wp"], L["bp"])
    gam, bet = torch.chunk(film, 2, dim=-1)
    zc = (1.0 + torch.tanh(gam)) * z + bet
    logits = F.linear(zc, L["wh"], L["bh"])
    loss = F.binary_cross_entropy_with_logits(logits, y)
    loss.backward()
    d = lambda t: t.to(DEV).contiguous()          # noqa: E731
    G = {k: d(v) for k, v in prm.items()}
    gap_g, demo_g, y_g = d(gap), d(demo), d(y)
    wpT = G["wp"].t().contiguous()
    e = lambda *s: torch.full(s, float("nan"), device=DEV)       # noqa: E731
    bz, bh1, bh2, bfilm, bzc = e(B, F_), e(B, H), e(B, H), e(B, 2 * F_), e(B, F_)
    lg, dl, dz, dfilm, dh2, dh1, dgap = e(B, NL), e(B, NL), e(B, F_), e(B, 2 * F_), e(B, H), e(B, H), e(B, Cin)
    lp = e(lib.ecgb200_head_loss_parts(B))
    check(lib.ecgb200_mm_head_fwd_bwd_f32(ptr(gap_g), ptr(demo_g), ptr(wpT), ptr(G["wp"]), ptr(G["bp"]), ptr(G["w0"]),
                                          ptr(G["b0"]), ptr(G["w2"]), ptr(G["b2"]), ptr(G["wf"]), ptr(G["bf"]), ptr(G["wh"]),
                                          ptr(G["bh"]), ptr(y_g), ptr(bz), ptr(bh1), ptr(bh2), ptr(bfilm), ptr(bzc), ptr(lg),
                                          ptr(dl), ptr(dz), ptr(dfilm), ptr(dh2), ptr(dh1), ptr(dgap), ptr(lp), B, Cin, F_,
                                          D, H, NL, 1.0, stream()), "mm_head")
    import ctypes as C
    V5, I5 = C.c_void_p * 5, C.c_int * 5
    dw = [e(F_, Cin), e(NL, F_), e(2 * F_, H), e(H, H), e(H, D)]
    db = [e(F_), e(NL), e(2 * F_), e(H), e(H)]
    ls = e(1)
    check(lib.ecgb200_head_wgrad_multi_f32(5, V5(ptr(dz), ptr(dl), ptr(dfilm), ptr(dh2), ptr(dh1)),
                                           V5(ptr(gap_g), ptr(bzc), ptr(bh2), ptr(bh1), ptr(demo_g)),
                                           V5(*[ptr(t) for t in dw]), V5(*[ptr(t) for t in db]),
                                           I5(F_, NL, 2 * F_, H, H), I5(Cin, F_, H, H, D), ptr(lp), ptr(ls), B, NL,
                                           stream()), "head_wgrad_multi")
    torch.cuda.synchronize()
    tol = 2e-5
    assert rel_inf(lg, logits) < tol and rel_inf(bzc, zc) < tol and rel_inf(bfilm, film) < tol
    assert abs(float(ls) - float(loss.detach())) < 1e-6 * max(1.0, abs(float(loss.detach())))
    assert rel_inf(dgap, gl.grad) < tol
    for (w, b), (kw, kb) in zip(zip(dw, db), (("wp", "bp"), ("wh", "bh"), ("wf", "bf"), ("w2", "b2"), ("w0", "b0"))):
        assert rel_inf(w, L[kw].grad) < tol, kw
        assert rel_inf(b, L[kb].grad) < tol, kb


@pytest.mark.parametrize("B,C,L", [(5, 256, 125), (2, 128, 31), (256, 256, 125), (64, 256, 625), (3, 64, 250)])
def test_last_block_bn_backward_in_one_pass_from_the_routing_summary(B, C, L):
    """Block 4 (pooled output -> time average): the forward pass leaves per (window, channel) the number of routed pool pairs
    and the sum of the raw conv outputs at the routed positions; the backward pass forms its two batch reductions from
    dgap * route and reads y once.  Against autograd (small shapes) and against the two-pass kernels (all shapes)."""
    y = (gen(B, C, L, seed=5) * 1.7 + 0.3).to(BF).float().requires_grad_(True)
    gamma = (1 + 0.2 * gen(C, seed=6)).requires_grad_(True)
    beta = (0.1 * gen(C, seed=7)).requires_grad_(True)
    yb = to_blocked(y.detach()).to(DEV)
    gg, bg = gamma.detach().to(DEV), beta.detach().to(DEV)
    yy = y.detach().double()
    part = torch.stack([yy.sum((0, 2)), (yy * yy).sum((0, 2))]).float().reshape(1, 2, C).to(DEV)
    st = torch.empty(4, C, device=DEV)
    gap = torch.empty(B, C, device=DEV)
    route = torch.full((2, B, C), float("nan"), device=DEV)
    check(lib.ecgb200_bn_relu_pool_fwd_train_route_bf16(ptr(yb), ptr(part), 1, ptr(gg), ptr(bg), None, None, None, ptr(st), None,
                                                        ptr(gap), ptr(route), B, C, L, 0.1, 1e-5, 1, stream()), "fwd_route")
    dout = gen(B, C, seed=10)
    dgap = dout.to(DEV)
    ws = torch.empty(lib.ecgb200_bn_bwd_ws_bytes(B, C), dtype=torch.uint8, device=DEV)
    ns = lib.ecgb200_bn_nsplit(B, C)
    res = []
    for one_pass in (False, True):
        dyb = torch.empty(B, C // 8, L, 8, dtype=BF, device=DEV)
        dgm, dbt = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
        dbp = torch.empty(C, ns, device=DEV)
        if one_pass:
            check(lib.ecgb200_bn_relu_pool_bwd_route_bf16(ptr(yb), ptr(st), ptr(dgap), ptr(route), ptr(dyb), ptr(dgm), ptr(dbt),
                                                          ptr(dbp), B, C, L, 1, stream()), "bwd_route")
        else:
            check(lib.ecgb200_bn_relu_pool_bwd_bf16(ptr(yb), ptr(st), None, ptr(dgap), ptr(dyb), ptr(dgm), ptr(dbt), ptr(dbp),
                                                    ptr(ws), B, C, L, 1, stream()), "bwd")
        torch.cuda.synchronize()
        res.append((from_blocked(dyb.cpu(), C), dgm.cpu(), dbt.cpu(), dbp.sum(1).cpu()))
    (dy2, dg2, db2, s2), (dy1, dg1, db1, s1) = res
    cnt = route[0].cpu()
    assert float(cnt.min()) >= 0 and float(cnt.max()) <= L // 2 and torch.equal(cnt, cnt.round())
    assert rel_inf(dg1, dg2) < 1e-5 and rel_inf(db1, db2) < 1e-5
    assert rel_inf(dy1, dy2) < 8e-3                                   # the same values up to one bf16 rounding
    assert (dy1 != dy2).float().mean().item() < 0.02
    if B <= 8:
        p_ref, out_ref, _, _, _ = None, None, None, None, None
        h = F.batch_norm(y, torch.zeros(C), torch.ones(C), gamma, beta, training=True, momentum=0.1, eps=1e-5)
        F.max_pool1d(F.relu(h), 2).mean(dim=2).backward(dout)
        assert rel_inf(dg1, gamma.grad) < 1e-4 and rel_inf(db1, beta.grad) < 1e-4
        assert rel_inf(dy1, y.grad) < 6e-3
