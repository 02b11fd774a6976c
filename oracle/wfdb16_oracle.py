"""CPU oracle for the input row N2 (TEST INFRASTRUCTURE ONLY -- see oracle/ecg_oracle.py's header).

Restates what the reference does to one record before the hot path:
  wfdb.rdsamp(path)              -> float64 physical signal (T, n_sig)        src/datasets/ptbxl.py:25
  np.asarray(sig, float32), .T   -> [12, T] float32                           ptbxl.py:29,36-50
  (x - mean) / (std + 1e-6)      per lead over time, numpy float32            ptbxl.py:122-127
`wfdb` (pinned in /root/reference/requirements.txt) is absent from this image and not vendored under
/root/reference, so the DECODE step is restated from the published WFDB format-16 specification
(little-endian int16 frames; physical = (digital - baseline) / gain; digital -32768 = NaN): parity of the decode
is unpinned; the normalisation lines are the reference's own numpy code, restated verbatim."""
import numpy as np


def rdsamp_format16(raw: bytes, gains, baselines) -> np.ndarray:
    n = len(gains)
    d = np.frombuffer(raw, dtype="<i2").reshape(-1, n).astype(np.int64)
    p = (d - np.asarray(baselines, dtype=np.int64)[None, :]).astype(np.float64) / np.asarray(gains, dtype=np.float64)[None, :]
    p[d == -32768] = np.nan
    return p                                        # (T, n_sig) float64, like wfdb.rdsamp()[0]


def load_and_normalize(raw: bytes, gains, baselines, normalize: bool = True) -> np.ndarray:
    sig = np.asarray(rdsamp_format16(raw, gains, baselines), dtype=np.float32)        # ptbxl.py:29
    x = sig.T                                                                          # ptbxl.py:36-50 -> [leads, T]
    if normalize:                                                                      # ptbxl.py:122-127
        mean = x.mean(axis=1, keepdims=True)
        std = x.std(axis=1, keepdims=True) + 1e-6
        x = (x - mean) / std
    return np.ascontiguousarray(x, dtype=np.float32)
