"""Evaluation helpers that follow the path: the threshold sweep of the fine-tuning CLIs and the best-F1 threshold
search of the BERT scripts, on the device.

    threshold_sweep(probs, labels, thresholds)            reference finetune_text.py:576-580 (sklearn P / R / F1)
    find_best_f1_and_threshold(scores, labels, ...)       reference finetune_bert.py:72-106
"""
import numpy as np
import torch

from . import _lib
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
    """(best_acc, best_f1, best_precision, best_recall, threshold) as reference finetune_bert.py:72-106 computes them with a
    Python sort + loop: stable sort by score, prefix counts, first maximum of F1 over the first n-1 cut points, threshold
    halfway to the next score.  One C-ABI call (ia_best_f1_threshold): radix sort + scan + F1 + arg-max kernels, float64
    arithmetic in the loop's own order; only the five results travel to the host."""
    if not scores.is_cuda:
        raise RuntimeError("item_alignment_b200 runs on CUDA tensors only (no CPU fallback)")
    scores = scores.detach().reshape(-1)
    if scores.dtype == torch.float64:
        dt = _lib.IA_F64
    else:
        scores, dt = scores.to(torch.float32), _lib.IA_F32        # bf16 / fp16 scores are exact in fp32
    scores = scores.contiguous()
    labels = labels.to(device=scores.device, dtype=torch.int64).reshape(-1).contiguous()
    n = scores.numel()
    assert n == labels.numel()
    if n < 2:
        return 0, 0, 0, 0, 0
    out = torch.empty(5, dtype=torch.float64, device=scores.device)
    with torch.cuda.device(scores.device):
        need = _lib.lib().ia_best_f1_workspace_bytes(n, dt)
        ws = torch.empty(need + 256, dtype=torch.uint8, device=scores.device)
        off = (-ws.data_ptr()) % 256
        _lib.check(_lib.lib().ia_best_f1_threshold(dt, scores.data_ptr(), labels.data_ptr(), n, int(bool(high_score_more_similar)),
                                                   out.data_ptr(), ws.data_ptr() + off, need, F_._stream()))
    acc, f1, precision, recall, threshold = out.tolist()
    if f1 <= 0:
        return 0, 0, 0, 0, 0
    return acc, f1, precision, recall, threshold
