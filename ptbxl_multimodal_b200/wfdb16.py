"""Input side of the hot path: WFDB format-16 records -> the (B, 12, T) float32, per-lead z-scored
tensor the reference's Datasets return (src/datasets/ptbxl.py:14-50,122-142), decoded ON THE DEVICE from the raw
.dat bytes.  The host only parses the tiny .hea text header and hands the bytes over (pinned, async H2D).

wfdb is a third-party dependency of the reference (requirements.txt) that is not vendored; the format restated here
is the published WFDB spec: header signal line `file fmt gain(baseline)/units adcres adczero initval checksum
blocksize description`, format 16 = little-endian two's-complement int16, sample-interleaved frames,
physical = (digital - baseline) / gain, digital -32768 = missing (NaN)."""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

from ._lib import lib, check, ptr, stream, EcgB200Error


@dataclass
class Wfdb16Header:
    n_sig: int
    fs: float
    n_samples: int
    gains: List[float]
    baselines: List[int]
    units: List[str]
    names: List[str]
    dat_files: List[str]


_SIG = re.compile(r"^(?P<gain>[-+0-9.eE]+)?(\((?P<base>-?\d+)\))?(/(?P<units>\S+))?$")


def parse_header(text: str) -> Wfdb16Header:
    """Parse the text of a .hea file whose signals are all format 16 in one .dat file (PTB-XL layout)."""
    lines = [ln.strip() for ln in text.splitlines() if ln.strip() and not ln.lstrip().startswith("#")]
    if not lines:
        raise EcgB200Error("empty WFDB header")
    rec = lines[0].split()
    if len(rec) < 2:
        raise EcgB200Error("malformed WFDB record line")
    n_sig = int(rec[1].split("/")[0])
    fs = float(rec[2].split("/")[0].split("(")[0]) if len(rec) > 2 else 250.0
    n_samples = int(rec[3]) if len(rec) > 3 else 0
    gains, baselines, units, names, files = [], [], [], [], []
    for ln in lines[1:1 + n_sig]:
        f = ln.split()
        if len(f) < 2:
            raise EcgB200Error(f"malformed WFDB signal line: {ln!r}")
        if f[1].split("x")[0].split(":")[0].split("+")[0] != "16":
            raise EcgB200Error(f"only WFDB format 16 is decoded on the device (got {f[1]!r})")
        m = _SIG.match(f[2]) if len(f) > 2 else None
        gain = float(m.group("gain")) if m and m.group("gain") else 200.0      # WFDB default gain
        if gain == 0:
            gain = 200.0
        adc_zero = int(f[4]) if len(f) > 4 else 0
        base = int(m.group("base")) if m and m.group("base") is not None else adc_zero
        gains.append(gain)
        baselines.append(base)
        units.append(m.group("units") if m and m.group("units") else "mV")
        names.append(" ".join(f[8:]) if len(f) > 8 else "")
        files.append(f[0])
    if len(gains) != n_sig:
        raise EcgB200Error("WFDB header lists fewer signals than it declares")
    return Wfdb16Header(n_sig, fs, n_samples, gains, baselines, units, names, files)


def decode_batch(frames: torch.Tensor, gains: Sequence[float], baselines: Sequence[int],
                 normalize: bool = True) -> torch.Tensor:
    """frames: CUDA int16 tensor (B, T, n_leads) of raw .dat frames -> (B, n_leads, T) float32 on the same device:
    physical units, transposed, per-lead z-scored when `normalize` (ptbxl.py:122-127)."""
    if not frames.is_cuda or frames.dtype != torch.int16 or frames.dim() != 3 or not frames.is_contiguous():
        raise EcgB200Error("decode_batch needs a contiguous CUDA int16 tensor (B, T, n_leads); no CPU fallback")
    b, t, n = frames.shape
    if len(gains) != n or len(baselines) != n:
        raise EcgB200Error("one gain and one baseline per lead")
    g = torch.tensor(list(gains), dtype=torch.float32, device=frames.device)
    bl = torch.tensor(list(baselines), dtype=torch.int32, device=frames.device)
    out = torch.empty(b, n, t, dtype=torch.float32, device=frames.device)
    check(lib.ecgb200_wfdb16_zscore_f32(ptr(frames), ptr(g), ptr(bl), ptr(out), b, n, t, 1 if normalize else 0, stream()),
          "wfdb16_zscore")
    return out


def frames_from_bytes(raw: bytes, n_leads: int) -> np.ndarray:
    """View the bytes of a format-16 .dat file as (T, n_leads) int16 frames (little endian)."""
    a = np.frombuffer(raw, dtype="<i2")
    if a.size % n_leads:
        raise EcgB200Error(".dat size is not a whole number of frames")
    return a.reshape(-1, n_leads)
