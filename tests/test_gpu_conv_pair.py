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
    lib.ecgb200_debug_set_conv_pair(1 if pair else 0)
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
        lib.ecgb200_debug_set_conv_pair(1)


SHAPES = [(1, 64, 128, 250), (3, 64, 128, 250), (7, 128, 256, 125), (2, 256, 128, 125), (5, 128, 64, 250),
          (256, 128, 256, 125), (256, 256, 128, 125), (256, 64, 128, 250), (256, 128, 64, 250),
          (301, 128, 256, 125), (37, 64, 128, 129), (2, 128, 256, 625), (64, 64, 128, 1250), (1, 256, 256, 40),
          (149, 128, 256, 125), (255, 128, 64, 250), (253, 64, 128, 250), (255, 256, 128, 125), (150, 256, 256, 125)]
PAIRED = {(256, 128, 256, 125), (256, 256, 128, 125), (256, 64, 128, 250), (256, 128, 64, 250), (301, 128, 256, 125),
          (64, 64, 128, 1250), (149, 128, 256, 125), (255, 128, 64, 250), (253, 64, 128, 250), (255, 256, 128, 125),
          (150, 256, 256, 125)}


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
    y1, s1, p1 = run_conv(xb, wf, bias, B, Ci, Co, L, False, stats)
    y2, s2, p2 = run_conv(xb, wf, bias, B, Ci, Co, L, True, stats)
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
    if B <= 8:
        refc = F.conv1d(x.to(BF).float(), w.to(BF).float(), bias.cpu(), padding=7)
        got = from_blocked(y2.cpu(), Co)
        assert float((got - refc).abs().max() / refc.abs().max()) <= 1e-2


def test_small_layers_keep_the_one_sm_kernel():
    # weights that fit in shared memory: nothing to halve, same partial count either way
    for (B, Ci, Co, L) in [(8, 16, 32, 1000), (8, 32, 64, 500), (8, 64, 32, 500)]:
        lib.ecgb200_debug_set_conv_pair(0)
        a = lib.ecgb200_conv1d_stat_parts_bf16(B, Ci, Co, L)
        lib.ecgb200_debug_set_conv_pair(1)
        assert lib.ecgb200_conv1d_stat_parts_bf16(B, Ci, Co, L) == a
