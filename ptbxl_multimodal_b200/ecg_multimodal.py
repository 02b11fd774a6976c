"""B200-native drop-in for the reference's ``src/models/ecg_multimodal.py``
(ECGBackbone :19-41, DemoEncoder :44-59, ECGMultimodal FiLM model :62-99).
Same constructors, attribute tree and ``state_dict`` keys; differentiable w.r.t.
the parameters and the demographic vector (scripts/12_grad_cam_ecg_demo.py:85-91)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as Fn
from .ecg_cnn import ConvBlock, _backbone_gap  # noqa: F401  (ConvBlock re-exported like the reference's duplicate)


class ECGBackbone(nn.Module):
    """Input [B, in_leads, T] -> [B, feat_dim]   (ecg_multimodal.py:19-41)."""

    def __init__(self, in_leads: int = 12, feat_dim: int = 256):
        super().__init__()
        chs = [32, 64, 128, 256]
        c = in_leads
        blocks = []
        for n in chs:
            blocks.append(ConvBlock(c, n))
            c = n
        self.backbone = nn.Sequential(*blocks)
        self.gap = nn.AdaptiveAvgPool1d(1)
        self.proj = nn.Linear(chs[-1], feat_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        g = _backbone_gap(self.backbone, x)
        return Fn.linear(g, self.proj.weight, self.proj.bias)


class DemoEncoder(nn.Module):
    """[age_norm, sex_id, height_norm, weight_norm, pacemaker] -> hidden (ecg_multimodal.py:44-59)."""

    def __init__(self, demo_dim: int = 5, hidden_dim: int = 64):
        super().__init__()
        self.mlp = nn.Sequential(
            nn.Linear(demo_dim, 64),
            nn.ReLU(inplace=True),
            nn.Linear(64, hidden_dim),
            nn.ReLU(inplace=True),
        )

    def forward(self, x_demo: torch.Tensor) -> torch.Tensor:
        h = Fn.linear(x_demo, self.mlp[0].weight, self.mlp[0].bias, act=1)
        return Fn.linear(h, self.mlp[2].weight, self.mlp[2].bias, act=1)


class ECGMultimodal(nn.Module):
    """FiLM-conditioned multimodal model (ecg_multimodal.py:62-99)."""

    def __init__(self, in_leads: int = 12, feat_dim: int = 256, demo_dim: int = 5, num_labels: int = 5,
                 demo_hidden_dim: int = 64, ecg_feat_dim: int = None, **kwargs):
        super().__init__()
        if ecg_feat_dim is not None:
            feat_dim = ecg_feat_dim
        self.ecg_backbone = ECGBackbone(in_leads=in_leads, feat_dim=feat_dim)
        self.demo_encoder = DemoEncoder(demo_dim=demo_dim, hidden_dim=demo_hidden_dim)
        self.film_gen = nn.Linear(demo_hidden_dim, 2 * feat_dim)
        self.head = nn.Linear(feat_dim, num_labels)

    def forward(self, x_ecg: torch.Tensor, x_demo: torch.Tensor) -> torch.Tensor:
        z_ecg = self.ecg_backbone(x_ecg)
        h_demo = self.demo_encoder(x_demo)
        film = Fn.linear(h_demo, self.film_gen.weight, self.film_gen.bias)
        z_cond = Fn.FilmFn.apply(z_ecg, film)          # (1 + tanh(gamma)) * z + beta
        return Fn.linear(z_cond, self.head.weight, self.head.bias)
