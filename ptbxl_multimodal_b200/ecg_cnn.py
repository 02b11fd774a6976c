"""B200-native drop-in for the reference's ``src/models/ecg_cnn.py``.

Same constructor signatures, attribute tree and ``state_dict`` keys as the
reference (ConvBlock at src/models/ecg_cnn.py:5-20, ECGCNN at :23-68), so the
shipped ``.pth`` checkpoints load with ``strict=True`` and callers that reach
``model.backbone[-1].net[0]`` (scripts/11_grad_cam_ecg_baseline.py:111) still find
an ``nn.Conv1d`` whose forward / backward hooks see the raw conv output.
Every operator underneath is a hand-written sm_100a kernel from libecgb200.so;
the ``nn.*`` sub-modules only hold parameters (and keep the default-init RNG
order, so a shared seed gives bit-identical random-init weights)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as Fn


class B200Conv1d(nn.Conv1d):
    """nn.Conv1d(k=15, padding=7) whose forward is the ecgb200 conv kernel.  In
    training it also emits the per-tile BatchNorm statistics from its epilogue."""

    _want_stats = False
    _stat = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if (self.kernel_size != (15,) or self.padding != (7,) or self.stride != (1,)
                or self.dilation != (1,) or self.groups != 1 or self.padding_mode != "zeros"):
            raise NotImplementedError("ecgb200 Conv1d implements kernel_size=15, padding=7, stride=1 only")
        y, stat = Fn.Conv1dK15Fn.apply(x, self.weight, self.bias, bool(self._want_stats))
        self._stat = (stat, y) if stat.numel() else None
        return y

    def take_stat(self, y: torch.Tensor):
        """Epilogue statistics, valid only for the very tensor this module just produced."""
        st, self._stat = self._stat, None
        if st is not None and st[1] is y:
            return st[0]
        return None


class ConvBlock(nn.Module):
    """Conv1d -> BatchNorm1d -> ReLU -> MaxPool1d(2)   (src/models/ecg_cnn.py:5-20)."""

    def __init__(self, in_ch: int, out_ch: int, k: int = 15, p: int = 2):
        super().__init__()
        if k != 15 or p != 2:
            raise NotImplementedError("ecgb200 ConvBlock implements k=15, p=2 (the reference's only use)")
        self.net = nn.Sequential(
            B200Conv1d(in_ch, out_ch, kernel_size=k, padding=k // 2),
            nn.BatchNorm1d(out_ch),
            nn.ReLU(inplace=True),
            nn.MaxPool1d(kernel_size=p),
        )
        self.emit_gap = False      # set on the last block: also produce mean over time
        self.gap_out = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        conv, bn = self.net[0], self.net[1]
        training = bn.training or not bn.track_running_stats
        conv._want_stats = training
        a = conv(x)                                   # module call: user hooks on the Conv1d fire
        if bn.momentum is None or not bn.affine or not bn.track_running_stats:
            raise NotImplementedError("ecgb200 BatchNorm1d implements the default (affine, momentum=0.1) config")
        p, gap = Fn.BnReluPoolFn.apply(a, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                       bn.num_batches_tracked, conv.take_stat(a), training,
                                       float(bn.momentum), float(bn.eps), self.emit_gap)
        self.gap_out = gap if self.emit_gap else None
        return p


def _backbone_gap(backbone: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    """backbone(x) followed by AdaptiveAvgPool1d(1)+squeeze, the mean fused into the last
    block's epilogue (ecg_cnn.py:61-62)."""
    last = backbone[-1]
    last.emit_gap = True
    try:
        h = backbone(x)
        g = last.gap_out
    finally:
        last.gap_out = None
    if g is None:                                     # a hook replaced the block output
        g = h.mean(dim=2)
    return g


class ECGCNN(nn.Module):
    """CNN encoder for 12-lead ECG classification (src/models/ecg_cnn.py:23-68)."""

    def __init__(self, in_leads: int = 12, feat_dim: int = 256, num_labels: int = 3):
        super().__init__()
        channels = [32, 64, 128, 256]
        c = in_leads
        blocks = []
        for n in channels:
            blocks.append(ConvBlock(c, n))
            c = n
        self.backbone = nn.Sequential(*blocks)
        self.gap = nn.AdaptiveAvgPool1d(1)            # kept for tree parity; fused into block 4
        self.proj = nn.Linear(channels[-1], feat_dim)
        self.head = nn.Linear(feat_dim, num_labels)

    def forward(self, x: torch.Tensor, return_features: bool = False):
        g = _backbone_gap(self.backbone, x)                           # [B, 256]
        z = Fn.linear(g, self.proj.weight, self.proj.bias)            # [B, feat_dim]
        logits = Fn.linear(z, self.head.weight, self.head.bias)       # [B, num_labels]
        if return_features:
            return logits, z
        return logits
