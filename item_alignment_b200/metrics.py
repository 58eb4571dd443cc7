"""Evaluation helpers that follow the path: the threshold sweep of the fine-tuning CLIs and the best-F1 threshold
search of the BERT scripts, on the device.

    threshold_sweep(probs, labels, thresholds)            reference finetune_text.py:576-580 (sklearn P / R / F1)
    find_best_f1_and_threshold(scores, labels, ...)       reference finetune_bert.py:72-106
"""
import numpy as np
import torch

from . import functional as F_


def threshold_sweep(probs, labels, thresholds=None):
    """precision, recall, f1 (float64 numpy arrays, one entry per threshold) of `probs >= threshold`, exactly what
    sklearn's precision_score / recall_score / f1_score return in the reference loop (0.0 where undefined).
    The confusion counts come from one CUDA kernel (ia_threshold_sweep); the ratios are formed from the integers."""
    if thresholds is None:
        thresholds = np.arange(0.1, 1.0, 0.1)                  # the reference's grid
    thresholds = [float(t) for t in thresholds]
    out_p, out_r, out_f = [], [], []
    for s in range(0, len(thresholds), 32):
        c = F_.threshold_sweep_counts(probs, labels, thresholds[s:s + 32]).cpu().numpy().astype(np.float64)
        tp, fp, fn = c[:, 0], c[:, 1], c[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            p = np.where(tp + fp > 0, tp / (tp + fp), 0.0)
            r = np.where(tp + fn > 0, tp / (tp + fn), 0.0)
            f = np.where(2 * tp + fp + fn > 0, 2 * tp / (2 * tp + fp + fn), 0.0)
        out_p.append(p); out_r.append(r); out_f.append(f)
    return np.concatenate(out_p), np.concatenate(out_r), np.concatenate(out_f)


def find_best_f1_and_threshold(scores, labels, high_score_more_similar: bool = True):
    """(best_acc, best_f1, best_precision, best_recall, threshold) as reference finetune_bert.py:72-106 computes them
    with a Python sort + loop: stable sort by score, prefix counts, first maximum of F1 over the first n-1 cut points,
    threshold halfway to the next score.  Done with device sort / scan primitives in float64 (same IEEE operations as
    the Python floats of the reference)."""
    if not scores.is_cuda:
        raise RuntimeError("item_alignment_b200 runs on CUDA tensors only (no CPU fallback)")
    scores = scores.detach().view(-1)
    labels = labels.to(scores.device).view(-1)
    n = scores.numel()
    assert n == labels.numel()
    if n < 2:
        return 0, 0, 0, 0, 0
    s64 = scores.to(torch.float64)
    order = torch.sort(s64, descending=high_score_more_similar, stable=True).indices
    ss = s64[order]
    ll = (labels[order] == 1).to(torch.float64)
    total_dup = float((labels.to(torch.float64)).sum())
    neg_total = n - total_dup
    ncorrect = torch.cumsum(ll, 0)[: n - 1]
    nextract = torch.arange(1, n, device=scores.device, dtype=torch.float64)
    fneg = nextract - ncorrect
    ok = ncorrect > 0
    precision = ncorrect / nextract
    recall = ncorrect / total_dup if total_dup > 0 else torch.zeros_like(ncorrect)
    f1 = torch.where(ok, 2 * precision * recall / (precision + recall), torch.zeros_like(precision))
    f1 = torch.nan_to_num(f1, nan=0.0)
    best = int(torch.argmax(f1))                               # first maximum, like the reference's strict '>'
    if float(f1[best]) <= 0:
        return 0, 0, 0, 0, 0
    acc = (ncorrect[best] + neg_total - fneg[best]) / n
    threshold = (ss[best] + ss[best + 1]) / 2
    return float(acc), float(f1[best]), float(precision[best]), float(recall[best]), float(threshold)
