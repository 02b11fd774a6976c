"""Loops for the ECG + demographics model: same contract as the reference's
src/training/loop_demo.py (train_one_epoch_demo :13-43, eval_one_epoch_demo :46-85);
epoch loss is the mean of batch means (loop_demo.py:38-43)."""
import numpy as np
import torch

from . import functional as Fn
from .metrics import compute_metrics


def train_one_epoch_demo(model, loader, optimizer, device, engine=None):
    """`engine`: optional ecgb200 TrainStep built for (model, optimizer, batch size); needs full batches (drop_last=True)."""
    model.train()
    total = torch.zeros((), dtype=torch.float64, device=device)
    num_batches = 0
    if engine is not None:
        for x_ecg, x_demo, y in loader:
            if x_ecg.shape[0] != engine.B:
                raise ValueError(f"TrainStep was built for batches of {engine.B} windows, the loader produced "
                                 f"{x_ecg.shape[0]}: use drop_last=True")
            total += engine(x_ecg, y, x_demo).double()
            num_batches += 1
        return float(total.item()) / max(1, num_batches)
    for x_ecg, x_demo, y in loader:
        x_ecg = x_ecg.to(device, non_blocking=True)
        x_demo = x_demo.to(device, non_blocking=True)
        y = y.to(device, non_blocking=True)
        optimizer.zero_grad()
        logits = model(x_ecg, x_demo)
        loss = Fn.binary_cross_entropy_with_logits(logits, y)
        loss.backward()
        optimizer.step()
        total += loss.detach().double()
        num_batches += 1
    return float(total.item()) / max(1, num_batches)


def eval_one_epoch_demo(model, loader, device, engine=None):
    """`engine`: optional ecgb200 InferStep built for this model (see loop.eval_one_epoch)."""
    model.eval()
    if engine is not None:
        engine.refresh()
    total = torch.zeros((), dtype=torch.float64, device=device)
    num_batches = 0
    all_probs, all_targets = [], []
    with torch.no_grad():
        for x_ecg, x_demo, y in loader:
            x_ecg = x_ecg.to(device, non_blocking=True)
            x_demo = x_demo.to(device, non_blocking=True)
            y = y.to(device, non_blocking=True)
            logits = engine(x_ecg, x_demo) if engine is not None else model(x_ecg, x_demo)
            total += Fn.binary_cross_entropy_with_logits(logits, y).double()
            num_batches += 1
            all_probs.append(Fn.sigmoid(logits))
            all_targets.append(y)
    y_true = torch.cat(all_targets).cpu().numpy()
    y_prob = torch.cat(all_probs).cpu().numpy()
    metrics = compute_metrics(y_true, y_prob)
    metrics["bce_loss"] = float(total.item()) / max(1, num_batches)
    return metrics
