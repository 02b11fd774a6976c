"""Stock-PyTorch restatement of the reference's module trees and loop bodies, for timing the
LIBRARY path (cuDNN / cuBLAS / ATen kernels) on the same B200 the ecgb200 kernels run on.

TEST / BENCH INFRASTRUCTURE ONLY (same rule as ecg_oracle.py): nothing under
``ptbxl_multimodal_b200/`` imports this; ``bench.py`` uses it for its ``gpu_reference`` leg and the
multi-GPU tests use it as the torch-DDP-semantics comparator.

What it restates (the reference cannot travel to the GPU box):
  * ``ConvBlock`` / ``ECGCNN``        /root/reference/src/models/ecg_cnn.py:5-68
  * ``ECGBackbone`` / ``DemoEncoder`` / ``ECGMultimodal``   src/models/ecg_multimodal.py:19-99
  * the loop body ``zero_grad -> model(x) -> BCE-with-logits -> backward -> optimizer.step``
    src/training/loop.py:22-36 and src/training/loop_demo.py:25-41, with ``torch.optim.AdamW``
    as the scripts construct it (scripts/03_train_ecg_baseline.py:133).
The modules are plain ``torch.nn`` layers with the reference's ``state_dict`` key names, so
``load_state_dict(ecg_oracle.init_state_dict(...))`` gives the oracle's weights; the CPU suite checks
that one step of this module equals ``ecg_oracle.train_step`` bit for bit.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Block(nn.Module):
    def __init__(self, cin: int, cout: int, k: int = 15):
        super().__init__()
        self.net = nn.Sequential(nn.Conv1d(cin, cout, k, padding=k // 2), nn.BatchNorm1d(cout),
                                 nn.ReLU(inplace=True), nn.MaxPool1d(2))

    def forward(self, x):
        return self.net(x)


def _backbone(in_leads: int):
    chans = [32, 64, 128, 256]
    blocks, c = [], in_leads
    for n in chans:
        blocks.append(_Block(c, n))
        c = n
    return nn.Sequential(*blocks), c


class StockECGCNN(nn.Module):
    def __init__(self, in_leads: int = 12, feat_dim: int = 256, num_labels: int = 3):
        super().__init__()
        self.backbone, c = _backbone(in_leads)
        self.gap = nn.AdaptiveAvgPool1d(1)
        self.proj = nn.Linear(c, feat_dim)
        self.head = nn.Linear(feat_dim, num_labels)

    def forward(self, x):
        z = self.proj(self.gap(self.backbone(x)).squeeze(-1))
        return self.head(z)


class _StockBackbone(nn.Module):
    def __init__(self, in_leads: int = 12, feat_dim: int = 256):
        super().__init__()
        self.backbone, c = _backbone(in_leads)
        self.gap = nn.AdaptiveAvgPool1d(1)
        self.proj = nn.Linear(c, feat_dim)

    def forward(self, x):
        return self.proj(self.gap(self.backbone(x)).squeeze(-1))


class _StockDemo(nn.Module):
    def __init__(self, in_dim: int = 5, hidden: int = 64):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(in_dim, hidden), nn.ReLU(inplace=True), nn.Linear(hidden, hidden),
                                 nn.ReLU(inplace=True))

    def forward(self, d):
        return self.mlp(d)


class StockECGMultimodal(nn.Module):
    def __init__(self, in_leads: int = 12, feat_dim: int = 256, demo_dim: int = 5, num_labels: int = 5,
                 demo_hidden_dim: int = 64):
        super().__init__()
        self.ecg_backbone = _StockBackbone(in_leads, feat_dim)
        self.demo_encoder = _StockDemo(demo_dim, demo_hidden_dim)
        self.film_gen = nn.Linear(demo_hidden_dim, 2 * feat_dim)
        self.head = nn.Linear(feat_dim, num_labels)

    def forward(self, x, d):
        z = self.ecg_backbone(x)
        g, b = torch.chunk(self.film_gen(self.demo_encoder(d)), 2, dim=-1)
        return self.head((1.0 + torch.tanh(g)) * z + b)


def build(kind: str, state_dict: Dict[str, torch.Tensor], num_labels: int = 5) -> nn.Module:
    m = StockECGCNN(12, 256, num_labels) if kind == "cnn" else StockECGMultimodal(num_labels=num_labels)
    m.load_state_dict({k: v.clone() for k, v in state_dict.items()}, strict=True)
    return m


def make_step(model: nn.Module, opt: torch.optim.Optimizer, autocast_dtype: Optional[torch.dtype] = None
              ) -> Callable[..., torch.Tensor]:
    """The reference loop body (loop.py:26-34) as a closure step(x, y[, demo]) -> loss tensor (no host sync)."""
    dev_type = next(model.parameters()).device.type

    def step(x, y, demo=None):
        opt.zero_grad(set_to_none=True)
        if autocast_dtype is not None:
            with torch.autocast(dev_type, dtype=autocast_dtype):
                out = model(x) if demo is None else model(x, demo)
            loss = F.binary_cross_entropy_with_logits(out.float(), y)
        else:
            out = model(x) if demo is None else model(x, demo)
            loss = F.binary_cross_entropy_with_logits(out, y)
        loss.backward()
        opt.step()
        return loss.detach()

    return step
