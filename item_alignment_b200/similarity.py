"""Competition scoring plugin with the reference's signature (submit/similarity.py:27):

    compute(item_emb_1: List[float], item_emb_2: List[float]) -> float

The organiser harness calls compute() once per pair on the CPU.  The reference's active body is
an ensemble pass-through (`return item_emb_2[0]`, :27-28); its commented body is the numpy softmax head
(:19-24) and pred_bert.py:47-52 is a pure-Python inner product.  `configure()` selects which of the reference's
measures compute() evaluates; anything other than the pass-through runs on the GPU through the HOST-buffer
C-ABI entry point (ia_pair_score_host).  `compute_many()` is the batched form a harness should prefer.
"""
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import MEASURES, check, lib

_cfg = {"measure": "passthrough", "w": None, "b": None, "device": 0}


def configure(measure: str = "passthrough", w=None, b=None, device: int = 0):
    if measure not in ("passthrough", "softmax") and measure not in MEASURES:
        raise ValueError(f"Unsupported similarty measure: {measure}")
    if measure == "softmax" and (w is None or b is None):
        raise ValueError("softmax needs the out_proj weight [2, 2D] and bias [2]")
    _cfg.update(measure=measure, w=None if w is None else np.asarray(w, dtype=np.float32),
                b=None if b is None else np.asarray(b, dtype=np.float32), device=device)


def compute_many(embs_1: Sequence[Sequence[float]], embs_2: Sequence[Sequence[float]], threshold: Optional[float] = None):
    """Scores (probabilities for softmax) of many pairs; with `threshold` also the `>= threshold` labels."""
    m = _cfg["measure"]
    e2 = np.ascontiguousarray(np.asarray(embs_2, dtype=np.float32))
    if m == "passthrough":
        s = e2[:, 0].astype(np.float64)
        return (s, s >= threshold) if threshold is not None else s
    e1 = np.ascontiguousarray(np.asarray(embs_1, dtype=np.float32))
    n, d = e1.shape
    if m == "softmax":
        import torch
        from . import functional as F_
        dev = torch.device("cuda", _cfg["device"])
        _, probs = F_.softmax_head(torch.from_numpy(e1).to(dev), torch.from_numpy(e2).to(dev),
                                   torch.from_numpy(_cfg["w"]).to(dev), torch.from_numpy(_cfg["b"]).to(dev))
        p = probs[:, 1].cpu().numpy()
        return (p, p.astype(np.float64) >= threshold) if threshold is not None else p
    sim = np.empty(n, dtype=np.float32)
    probs = np.empty(n, dtype=np.float32)
    labels = np.empty(n, dtype=np.uint8) if threshold is not None else None
    check(lib().ia_pair_score_host(MEASURES[m], _lib.IA_F32, e1.ctypes.data, e2.ctypes.data, n, d, sim.ctypes.data,
                                   probs.ctypes.data, float(threshold) if threshold is not None else 0.0,
                                   labels.ctypes.data if labels is not None else None, _cfg["device"]))
    return (sim, labels.astype(bool)) if threshold is not None else sim


def compute(item_emb_1: List[float], item_emb_2: List[float]) -> float:
    if _cfg["measure"] == "passthrough":
        return item_emb_2[0]                       # reference submit/similarity.py:27-28
    return float(compute_many([item_emb_1], [item_emb_2])[0])
