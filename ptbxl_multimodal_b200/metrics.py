"""Macro AUROC / AUPRC / F1 (CPU post-processing, once per epoch; out of the hot path).
Same contract as the reference's src/training/metrics.py:5-42."""
import numpy as np
from sklearn.metrics import average_precision_score, f1_score, roc_auc_score


def compute_metrics(y_true: np.ndarray, y_prob: np.ndarray, threshold: float = 0.5):
    metrics = {}
    for key, fn in (("auroc_macro", roc_auc_score), ("auprc_macro", average_precision_score)):
        try:
            metrics[key] = fn(y_true, y_prob, average="macro")
        except ValueError:
            metrics[key] = float("nan")
    y_pred = (y_prob >= threshold).astype(int)
    metrics["f1_macro"] = f1_score(y_true, y_pred, average="macro", zero_division=0)
    return metrics


def f1_macro_from_counts(counts) -> float:
    """Macro F1 from per-label confusion counts (tp, fp, fn, tn) as accumulated on the device by
    functional.eval_counts -- equals sklearn's f1_score(average="macro", zero_division=0) used at
    src/training/metrics.py:38-40."""
    c = np.asarray(counts, dtype=np.float64).reshape(-1, 4)
    tp, fp, fn = c[:, 0], c[:, 1], c[:, 2]
    den = 2 * tp + fp + fn
    f1 = np.where(den > 0, 2 * tp / np.maximum(den, 1), 0.0)
    return float(f1.mean())
