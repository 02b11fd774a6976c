"""Input row N2: WFDB format-16 header parsing (CPU) and the device decode + z-score against the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import wfdb16_oracle as W

HEA = """00001_hr 12 500 5000
00001_hr.dat 16 1000.0(0)/mV 16 0 -115 13047 0 I
00001_hr.dat 16 1000.0(0)/mV 16 0 -50 11561 0 II
00001_hr.dat 16 1000.0(0)/mV 16 0 65 64050 0 III
00001_hr.dat 16 1000.0(0)/mV 16 0 82 20510 0 AVR
00001_hr.dat 16 1000.0(0)/mV 16 0 -90 7302 0 AVL
00001_hr.dat 16 1000.0(0)/mV 16 0 7 5270 0 AVF
00001_hr.dat 16 1000.0(0)/mV 16 0 -65 21229 0 V1
00001_hr.dat 16 1000.0(0)/mV 16 0 -40 6400 0 V2
00001_hr.dat 16 1000.0(0)/mV 16 0 -5 21794 0 V3
00001_hr.dat 16 500.0(-12)/mV 16 0 -35 26713 0 V4
00001_hr.dat 16 1000.0/mV 16 3 -35 26713 0 V5
00001_hr.dat 16 1000.0(0)/mV 16 0 -75 14210 0 V6
# a comment
"""


def test_parse_header():
    from ptbxl_multimodal_b200.wfdb16 import parse_header, frames_from_bytes
    h = parse_header(HEA)
    assert h.n_sig == 12 and h.fs == 500 and h.n_samples == 5000
    assert h.gains[0] == 1000.0 and h.gains[9] == 500.0 and h.baselines[9] == -12
    assert h.baselines[10] == 3                      # no (baseline): falls back to adc_zero
    assert h.names[3] == "AVR" and h.units[0] == "mV" and set(h.dat_files) == {"00001_hr.dat"}
    raw = np.arange(24, dtype="<i2").tobytes()
    assert frames_from_bytes(raw, 12).shape == (2, 12)
    with pytest.raises(Exception):
        parse_header(HEA.replace(" 16 1000.0(0)/mV 16 0 -115", " 212 1000.0(0)/mV 12 0 -115"))


def test_oracle_decode_known_values():
    raw = np.array([[100, -32768], [300, 50], [500, 150]], dtype="<i2").tobytes()
    p = W.rdsamp_format16(raw, [200.0, 100.0], [100, -50])
    assert np.allclose(p[:, 0], [0.0, 1.0, 2.0]) and np.isnan(p[0, 1]) and np.allclose(p[1:, 1], [1.0, 2.0])
    x = W.load_and_normalize(np.array([[1], [2], [3], [4]], dtype="<i2").tobytes(), [1.0], [0])
    assert x.shape == (1, 4) and abs(float(x.mean())) < 1e-6 and abs(float(x.std()) - 1.0) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("B,T", [(3, 5000), (2, 1000), (1, 37)])
def test_device_decode_matches_oracle(B, T):
    from ptbxl_multimodal_b200.wfdb16 import parse_header, decode_batch
    h = parse_header(HEA)
    rng = np.random.default_rng(0)
    frames = (rng.standard_normal((B, T, 12)) * 300 + rng.integers(-200, 200, size=(1, 1, 12))).astype("<i2")
    ref = np.stack([W.load_and_normalize(frames[b].tobytes(), h.gains, h.baselines) for b in range(B)])
    out = decode_batch(torch.from_numpy(frames.astype(np.int16)).cuda(), h.gains, h.baselines).cpu().numpy()
    assert out.shape == (B, 12, T)
    assert np.abs(out - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())       # fp32 vs numpy's fp32 pairwise sums
    phys = decode_batch(torch.from_numpy(frames.astype(np.int16)).cuda(), h.gains, h.baselines, normalize=False).cpu().numpy()
    refp = np.stack([W.load_and_normalize(frames[b].tobytes(), h.gains, h.baselines, normalize=False) for b in range(B)])
    assert np.array_equal(phys, refp)                                          # the decode itself is bit-exact
    # a missing sample (-32768) poisons only its own lead, exactly as numpy does in the reference pipeline
    frames[0, 5, 2] = -32768
    out2 = decode_batch(torch.from_numpy(frames.astype(np.int16)).cuda(), h.gains, h.baselines).cpu().numpy()
    assert np.isnan(out2[0, 2]).all() and np.isfinite(out2[0, 1]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("B,T", [(3, 5000), (4, 1000), (1, 38)])
def test_fused_decode_zscore_pack_matches_oracle_and_feeds_the_engine(B, T):
    """N1 + N2 fused (ecgb200_wfdb16_zscore_pack_bf16): raw frames -> z-scored bf16 in the first conv's blocked layout ==
    the numpy oracle rounded to bf16 (<= 1 bf16 ulp on a vanishing fraction: the mean / std sums are fp64 in another
    order), padding leads exactly zero; and a raw_input TrainStep fed the frames takes the same step as the fp32-input
    engine fed the decoded windows."""
    from ptbxl_multimodal_b200._lib import lib, check, ptr, stream
    from ptbxl_multimodal_b200.wfdb16 import parse_header, decode_batch
    h = parse_header(HEA)
    rng = np.random.default_rng(1)
    frames = (rng.standard_normal((B, T, 12)) * 300 + rng.integers(-200, 200, size=(1, 1, 12))).astype("<i2")
    ft = torch.from_numpy(frames.astype(np.int16)).cuda()
    gain = torch.tensor(h.gains, dtype=torch.float32, device="cuda")
    base = torch.tensor(h.baselines, dtype=torch.int32, device="cuda")
    xb = torch.full((B, 2, T, 8), float("nan"), dtype=torch.bfloat16, device="cuda")
    check(lib.ecgb200_wfdb16_zscore_pack_bf16(ptr(ft), ptr(gain), ptr(base), ptr(xb), B, 12, T, stream()), "decode_pack")
    torch.cuda.synchronize()
    got = xb.float().permute(0, 1, 3, 2).reshape(B, 16, T).cpu()
    assert (got[:, 12:] == 0).all()
    ref = torch.from_numpy(np.stack([W.load_and_normalize(frames[b].tobytes(), h.gains, h.baselines) for b in range(B)]))
    refb = ref.to(torch.bfloat16).float()
    diff = (got[:, :12] - refb).abs()
    assert float(diff.max()) <= 2.0 ** -7 * float(refb.abs().max())              # at most one bf16 ulp
    assert float((diff > 0).float().mean()) < 1e-3
    if T < 1000:
        return
    import ptbxl_multimodal_b200 as P
    from ptbxl_multimodal_b200.step import TrainStep
    y = (torch.rand(B, 5, generator=torch.Generator().manual_seed(2)) < 0.3).float().cuda()
    res = []
    for raw in (False, True):
        torch.manual_seed(42)
        m = P.ECGCNN(12, 256, 5).cuda().train()
        eng = TrainStep(m, P.FusedAdamW(m.parameters(), lr=1e-3, weight_decay=1e-4), B, T, precision="bf16", raw_input=raw)
        if raw:
            eng.set_calibration(h.gains, h.baselines)
            eng.load_frames(ft, y)
            loss = float(eng.run())
        else:
            loss = float(eng(decode_batch(ft, h.gains, h.baselines), y))
        res.append((loss, eng.P.clone()))
    assert abs(res[0][0] - res[1][0]) < 1e-3 * abs(res[0][0])
    assert float((res[0][1] - res[1][1]).abs().max()) < 5e-3                     # a handful of 1-ulp input differences
