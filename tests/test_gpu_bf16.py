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
    check(lib.ecgb200_conv1d_fwd_bf16(ptr(xb), ptr(wf), ptr(bias.to(DEV)), ptr(yb), B, Ci, Co, L, stream()), "conv_tc")
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
    check(lib.ecgb200_conv1d_fwd_bf16(ptr(to_blocked(dy).to(DEV)), ptr(wd), None, ptr(dxb), B, Co, Ci, L, stream()), "dgrad_tc")
    torch.cuda.synchronize()
    assert rel_inf(from_blocked(dxb.cpu(), Ci), xr_.grad) < 1e-2
