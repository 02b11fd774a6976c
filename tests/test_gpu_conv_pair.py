"""GPU: the two-SM (tcgen05 cta_group::2) kernel of the streamed-weight conv layers against the one-SM kernel (bit-identical:
same operands, same order of the K steps per output) and against fp32 math on the bf16-rounded operands (rel_inf <= 1e-2).
Shapes: blocks 3 and 4 of src/models/ecg_cnn.py:29-33, forward (with the BatchNorm partial statistics) and dgrad, at batches that
give odd tile counts, a single pair, ragged last groups, one and two groups per pair."""
import pytest
import torch
import torch.nn.functional as F

from ptbxl_multimodal_b200._lib import lib, check, ptr, stream

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF = torch.bfloat16


def gen(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def to_blocked(x):
    b, c, l = x.shape
    return x.reshape(b, c // 8, 8, l).permute(0, 1, 3, 2).contiguous().to(BF)


def from_blocked(xb, c):
    b, cc, l, _ = xb.shape
    return xb.float().permute(0, 1, 3, 2).reshape(b, c, l)


def run_conv(xb, wf, bias, B, Ci, Co, L, pair, stats):
    lib.ecgb200_debug_set_conv_pair(pair)
    try:
        parts = lib.ecgb200_conv1d_stat_parts_bf16(B, Ci, Co, L)
        assert parts > 0
        yb = torch.full((B, Co // 8, L, 8), float("nan"), dtype=BF, device=DEV)
        sp = torch.zeros(parts, 2, Co, device=DEV) if stats else None
        check(lib.ecgb200_conv1d_fwd_stats_bf16(ptr(xb), ptr(wf), ptr(bias), ptr(yb), ptr(sp) if stats else None,
                                                B, Ci, Co, L, stream()), "conv")
        torch.cuda.synchronize()
        return yb, sp, parts
    finally:
        lib.ecgb200_debug_set_conv_pair(3)


SHAPES = [(1, 64, 128, 250), (3, 64, 128, 250), (7, 128, 256, 125), (2, 256, 128, 125), (5, 128, 64, 250),
          (256, 128, 256, 125), (256, 256, 128, 125), (256, 64, 128, 250), (256, 128, 64, 250),
          (301, 128, 256, 125), (37, 64, 128, 129), (2, 128, 256, 625), (64, 64, 128, 1250), (1, 256, 256, 40),
          (149, 128, 256, 125), (255, 128, 64, 250), (253, 64, 128, 250), (255, 256, 128, 125), (150, 256, 256, 125)]
PAIRED = {(256, 128, 256, 125), (256, 256, 128, 125), (256, 64, 128, 250), (256, 128, 64, 250), (301, 128, 256, 125),
          (64, 64, 128, 1250), (149, 128, 256, 125), (255, 128, 64, 250), (253, 64, 128, 250), (255, 256, 128, 125)}


@pytest.mark.parametrize("B,Ci,Co,L", SHAPES)
@pytest.mark.parametrize("stats", [True, False])
def test_pair_kernel_equals_one_sm_kernel(B, Ci, Co, L, stats):
    x = gen(B, Ci, L, seed=2)
    w = gen(Co, Ci, 15, seed=3, scale=0.05)
    bias = gen(Co, seed=4, scale=0.1).to(DEV)
    xb = to_blocked(x).to(DEV)
    wf = torch.empty(15, Ci // 8, Co, 8, dtype=BF, device=DEV)
    wd = torch.empty(15, Co // 8, Ci, 8, dtype=BF, device=DEV)
    check(lib.ecgb200_conv1d_prep_weights_bf16(ptr(w.to(DEV)), ptr(wf), ptr(wd), Co, Ci, stream()), "prep")
    y1, s1, p1 = run_conv(xb, wf, bias, B, Ci, Co, L, 0, stats)
    y2, s2, p2 = run_conv(xb, wf, bias, B, Ci, Co, L, 3, stats)
    assert p2 <= 148
    if (B, Ci, Co, L) in PAIRED:       # shapes the pair kernel takes on a 148-SM part (>= 2 tiles on the busiest CTA)
        assert p2 % 2 == 0, (p1, p2)
    else:                              # one tile per CTA or less: stays with the one-SM kernel
        assert p2 == p1
    assert not torch.isnan(y2.float()).any()
    assert torch.equal(y1, y2), f"max diff {(y1.float() - y2.float()).abs().max().item()}"
    if stats:
        # partials are grouped differently; their totals are sums of the same rounded outputs
        t1, t2 = s1.double().sum(0), s2.double().sum(0)
        yy = from_blocked(y2, Co).double()
        ref = torch.stack([yy.sum((0, 2)), (yy * yy).sum((0, 2))])
        assert (t2 - ref).abs().max() <= 1e-4 * ref.abs().max()
        assert (t1 - t2).abs().max() <= 1e-4 * ref.abs().max()
    # forced: the pair kernel also where one tile per SM makes it the slower choice
    y3, s3, p3 = run_conv(xb, wf, bias, B, Ci, Co, L, 7, stats)
    assert p3 % 2 == 0 and torch.equal(y1, y3)
    if stats:
        assert (s3.double().sum(0) - t2).abs().max() <= 1e-4 * ref.abs().max()
    if B <= 8:
        refc = F.conv1d(x.to(BF).float(), w.to(BF).float(), bias.cpu(), padding=7)
        got = from_blocked(y2.cpu(), Co)
        assert float((got - refc).abs().max() / refc.abs().max()) <= 1e-2


def test_small_layers_keep_the_one_sm_kernel():
    # weights that fit in shared memory: nothing to halve, same partial count either way
    for (B, Ci, Co, L) in [(8, 16, 32, 1000), (8, 32, 64, 500), (8, 64, 32, 500)]:
        lib.ecgb200_debug_set_conv_pair(0)
        a = lib.ecgb200_conv1d_stat_parts_bf16(B, Ci, Co, L)
        lib.ecgb200_debug_set_conv_pair(3)
        assert lib.ecgb200_conv1d_stat_parts_bf16(B, Ci, Co, L) == a


WG_SHAPES = [(5, 64, 128, 250), (6, 128, 256, 125), (2, 128, 256, 625), (256, 64, 128, 250), (256, 128, 256, 125),
             (77, 128, 256, 125), (3, 256, 256, 125), (9, 128, 128, 250), (1, 64, 128, 40)]


@pytest.mark.parametrize("B,Ci,Co,L", WG_SHAPES)
def test_pair_wgrad_equals_one_sm_wgrad(B, Ci, Co, L):
    """Blocks 3 / 4 weight gradient: the CTA-pair taps-as-M kernel against the one-SM taps-as-N kernel (same operands, fp32
    accumulation: rel_inf <= 1e-5 of each other) and, for small batches, against autograd on the bf16-rounded operands."""
    x = gen(B, Ci, L, seed=6)
    dy = gen(B, Co, L, seed=7)
    dyb, xb = to_blocked(dy).to(DEV), to_blocked(x).to(DEV)
    dbp = dy.to(BF).float().sum(dim=2).t().contiguous().to(DEV)
    out = []
    for pair in (0, 7):
        lib.ecgb200_debug_set_conv_pair(pair)
        try:
            ws = torch.empty(lib.ecgb200_conv1d_wgrad_bf16_ws_bytes(B, Ci, Co, L), dtype=torch.uint8, device=DEV)
            dw = torch.full((Co, Ci, 15), float("nan"), device=DEV)
            db = torch.full((Co,), float("nan"), device=DEV)
            check(lib.ecgb200_conv1d_wgrad_bf16(ptr(dyb), ptr(xb), ptr(dw), ptr(db), ptr(dbp), B, ptr(ws), B, Ci, Co, L,
                                                stream()), "wgrad")
            torch.cuda.synchronize()
            out.append((dw.cpu(), db.cpu()))
        finally:
            lib.ecgb200_debug_set_conv_pair(3)
    (dw0, db0), (dw1, db1) = out
    assert not torch.isnan(dw1).any()
    assert float((dw0 - dw1).abs().max() / dw0.abs().max()) <= 1e-5
    assert torch.equal(db0, db1)
    if B <= 9:
        w = torch.zeros(Co, Ci, 15, requires_grad=True)
        F.conv1d(x.to(BF).float(), w, None, padding=7).backward(dy.to(BF).float())
        assert float((dw1 - w.grad).abs().max() / w.grad.abs().max()) < 2e-3


@pytest.mark.parametrize("B,Ci,Co,L", [(256, 64, 128, 250), (256, 128, 256, 125), (77, 128, 256, 125), (3, 64, 128, 250),
                                       (512, 64, 128, 1250)])
@pytest.mark.parametrize("last", [False, True])
def test_pair_inference_blocks_equal_one_sm(B, Ci, Co, L, last):
    """Inference blocks (folded BatchNorm + ReLU + MaxPool in the conv epilogue; last block: time sums only) on the pair kernel:
    bit-identical to the one-SM kernel (forced on the small shapes)."""
    x = gen(B, Ci, L, seed=2)
    w = gen(Co, Ci, 15, seed=3, scale=0.05)
    scale = (1 + 0.2 * gen(Co, seed=4)).to(DEV)
    shift = (0.1 * gen(Co, seed=5)).to(DEV)
    xb = to_blocked(x).to(DEV)
    wf = torch.empty(15, Ci // 8, Co, 8, dtype=BF, device=DEV)
    wd = torch.empty(15, Co // 8, Ci, 8, dtype=BF, device=DEV)
    check(lib.ecgb200_conv1d_prep_weights_bf16(ptr(w.to(DEV)), ptr(wf), ptr(wd), Co, Ci, stream()), "prep")
    tiles = B * ((L + 127) // 128)
    out = []
    for mask in (0, 7):
        lib.ecgb200_debug_set_conv_pair(mask)
        try:
            pb = None if last else torch.full((B, Co // 8, L // 2, 8), float("nan"), dtype=BF, device=DEV)
            gp = torch.full((tiles, 4, Co), float("nan"), device=DEV) if last else None
            check(lib.ecgb200_conv1d_bn_relu_pool_infer_bf16(ptr(xb), ptr(wf), ptr(scale), ptr(shift), ptr(pb) if pb is not None else None,
                                                             ptr(gp) if gp is not None else None, B, Ci, Co, L, stream()), "infer")
            torch.cuda.synchronize()
            out.append(gp if last else pb)
        finally:
            lib.ecgb200_debug_set_conv_pair(3)
    assert not torch.isnan(out[1].float()).any()
    assert torch.equal(out[0], out[1])
