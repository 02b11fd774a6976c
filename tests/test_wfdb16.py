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
