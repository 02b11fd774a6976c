"""Training / evaluation loops for single-input ECG models: same signatures and return
values as the reference's src/training/loop.py (train_one_epoch :14-38,
eval_one_epoch :41-73), with the BCE / sigmoid running as ecgb200 kernels and the
per-step ``loss.item()`` host sync deferred to one read per epoch."""
from typing import Dict

import numpy as np
import torch

from . import functional as Fn
from .metrics import compute_metrics, f1_macro_from_counts


def _logits(out):
    return out[0] if isinstance(out, tuple) else out


def train_one_epoch(model, loader, optimizer, device, engine=None) -> float:
    """`engine`: optional ecgb200 TrainStep built for (model, optimizer, batch size): every step is then one CUDA-graph
    replay (zero_grad, forward, BCE, backward, AdamW).  The engine has static shapes: the loader must yield full batches
    (drop_last=True)."""
    model.train()
    total = torch.zeros((), dtype=torch.float64, device=device)
    if engine is not None:
        n = 0
        for x, y in loader:
            if x.shape[0] != engine.B:
                raise ValueError(f"TrainStep was built for batches of {engine.B} windows, the loader produced {x.shape[0]}: "
                                 "use drop_last=True")
            total += engine(x, y).double() * x.size(0)            # device scalar: no per-step sync
            n += x.size(0)
        return float(total.item()) / max(n, 1)
    for x, y in loader:
        x = x.to(device, non_blocking=True)
        y = y.to(device, non_blocking=True)
        optimizer.zero_grad()
        logits = _logits(model(x))
        loss = Fn.binary_cross_entropy_with_logits(logits, y)
        loss.backward()
        optimizer.step()
        total += loss.detach().double() * x.size(0)         # stays on device: no per-step sync
    return float(total.item()) / len(loader.dataset)


def eval_one_epoch(model, loader, device, engine=None) -> Dict[str, float]:
    """`engine`: optional ecgb200 InferStep built for this model (bf16 tensor-core forward as one CUDA graph);
    without it the fp32-exact module forward runs, as in the reference."""
    model.eval()
    if engine is not None:
        engine.refresh()
    all_targets, all_probs = [], []
    total = torch.zeros((), dtype=torch.float64, device=device)
    counts = None
    with torch.no_grad():
        for x, y in loader:
            x = x.to(device, non_blocking=True)
            y = y.to(device, non_blocking=True)
            logits = engine(x) if engine is not None else _logits(model(x))
            loss = Fn.binary_cross_entropy_with_logits(logits, y)
            total += loss.double() * x.size(0)
            if counts is None:
                counts = torch.zeros(logits.shape[1], 4, dtype=torch.int32, device=logits.device)
            prob, _ = Fn.eval_counts(logits, y, counts, threshold=0.5)      # sigmoid + threshold + confusion counts
            all_targets.append(y)
            all_probs.append(prob)
    y_true = torch.cat(all_targets).cpu().numpy()
    y_prob = torch.cat(all_probs).cpu().numpy()
    metrics = compute_metrics(y_true, y_prob, threshold=0.5)
    # thresholded metric from the device-side counts (no per-batch sync); identical to sklearn's value above
    metrics["f1_macro"] = f1_macro_from_counts(counts.cpu().numpy())
    metrics["bce_loss"] = float(total.item()) / len(loader.dataset)
    return metrics
