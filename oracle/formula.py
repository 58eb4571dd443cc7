"""Explicit numpy formulas for every score, probability map, loss and gradient on the path
-- TEST INFRASTRUCTURE (see oracle/__init__.py).

This layer does NOT call torch: it writes out the arithmetic that the reference delegates
to torch, including each op's eps convention, so that the CUDA
kernels, the torch port and the real reference modules can be triangulated.  Computed in
float64 by default (conditioning diagnostics) or in the dtype given.
"""
import numpy as np

COS_EPS = 1e-8        # nn.CosineSimilarity eps, reference src/models/base.py:58
PDIST_EPS = 1e-6      # nn.PairwiseDistance eps (added to the difference), base.py:60,62
COSEMB_EPS = 1e-12    # ATen cosine_embedding_loss EPSILON, reference text.py:1401


def _f(a, dt):
    return np.asarray(a, dtype=dt)


# ------------------------------------------------------------------ forward scores
def score(measure, x, y, dt=np.float64):
    """a1-a4: inner (base.py:29-34), cosine (base.py:58), l1/l2 (base.py:60-62)."""
    x, y = _f(x, dt), _f(y, dt)
    if measure == "inner_product":
        return (x * y).sum(1)
    if measure == "cosine":
        nx = np.maximum(np.sqrt((x * x).sum(1)), dt(COS_EPS))
        ny = np.maximum(np.sqrt((y * y).sum(1)), dt(COS_EPS))
        return ((x / nx[:, None]) * (y / ny[:, None])).sum(1)
    d = x - y + dt(PDIST_EPS)
    if measure == "l1":
        return np.abs(d).sum(1)
    if measure == "l2":
        return np.sqrt((d * d).sum(1))
    raise ValueError(f"Unsupported similarty measure: {measure}")


def probs(measure, s):
    """a5: base.py:79-86."""
    if measure == "cosine":
        return (s + 1) / 2
    if measure in ("l1", "l2"):
        return np.exp(-s)
    if measure == "inner_product":
        return 1.0 / (1.0 + np.exp(-s))
    raise ValueError(f"Unsupported similarty measure: {measure}")


# ------------------------------------------------------------------ d score / d x, y
def score_grad(measure, x, y, g, dt=np.float64):
    """a16: dx_i = g_i * d s_i / d x_i (and dy), g = dL/ds per pair."""
    x, y, g = _f(x, dt), _f(y, dt), _f(g, dt)[:, None]
    if measure == "inner_product":
        return g * y, g * x
    if measure == "cosine":
        # torch >= 2.0 (ATen cosine_similarity): n = clamp_min(norm, eps) applied in place under
        # no_grad, s = sum((x/n1)(y/n2)); autograd still differentiates n as the norm, so
        #   ds/dx = (yh - s * x/|x|) / n1      (x/|x| := 0 at |x| = 0)
        # which is the textbook (yh - s xh)/|x| whenever |x| >= eps.
        rx = np.sqrt((x * x).sum(1))[:, None]
        ry = np.sqrt((y * y).sum(1))[:, None]
        nx, ny = np.maximum(rx, COS_EPS), np.maximum(ry, COS_EPS)
        xh, yh = x / nx, y / ny
        s = (xh * yh).sum(1)[:, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            ux = np.where(rx > 0, x / rx, 0.0)
            uy = np.where(ry > 0, y / ry, 0.0)
        return g * (yh - s * ux) / nx, g * (xh - s * uy) / ny
    d = x - y + dt(PDIST_EPS)
    if measure == "l1":
        sg = np.sign(d)
        return g * sg, -g * sg
    if measure == "l2":
        s = np.sqrt((d * d).sum(1))[:, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = np.where(s > 0, d / s, 0.0)
        return g * u, -g * u
    raise ValueError(measure)


# ------------------------------------------------------------------ losses on a scalar score
def loss_per_pair(loss_type, s, labels, margin=1.0):
    """a8-a10 on sim (ladder text.py:1468-1477): labels in {0,1}; t = 2l-1.
    Returns (loss_i, dloss_i/ds_i)."""
    s = np.asarray(s, dtype=np.float64)
    l = np.asarray(labels, dtype=np.float64)
    t = 2 * l - 1
    if loss_type == "bce":       # nn.BCEWithLogitsLoss, text.py:1403,1477
        li = np.maximum(s, 0) - s * l + np.log1p(np.exp(-np.abs(s)))
        gi = 1.0 / (1.0 + np.exp(-s)) - l
    elif loss_type == "hinge":   # loss.py:126-134 ; sub-gradient 1/2 at the kink
        z = margin - s * t
        li = np.maximum(0.0, z)
        gi = np.where(z > 0, -t, np.where(z == 0, -0.5 * t, 0.0))
    elif loss_type == "euclidean":   # loss.py:61-68 ; s**t
        with np.errstate(divide="ignore"):
            li = np.where(t > 0, s, 1.0 / s)
            gi = np.where(t > 0, 1.0, -1.0 / (s * s))
    else:
        raise ValueError(loss_type)
    return li, gi


def cosine_embedding(x, y, labels, margin=1.0, dt=np.float64):
    """a11: nn.CosineEmbeddingLoss on the embeddings (text.py:1401,1471).
    c = sum(xy)/sqrt((sum(x^2)+1e-12)(sum(y^2)+1e-12)); t=1: 1-c ; t=-1: max(0,c-margin).
    Returns (loss_i, dx_i, dy_i) per pair before the mean."""
    x, y = _f(x, dt), _f(y, dt)
    t = 2 * np.asarray(labels, dtype=np.float64) - 1
    xy = (x * y).sum(1)
    a = (x * x).sum(1) + COSEMB_EPS
    b = (y * y).sum(1) + COSEMB_EPS
    den = np.sqrt(a * b)
    c = xy / den
    li = np.where(t > 0, 1 - c, np.maximum(0.0, c - margin))
    gc = np.where(t > 0, -1.0, np.where(c - margin > 0, 1.0, 0.0))     # dl/dc
    dcdx = y / den[:, None] - (c / a)[:, None] * x
    dcdy = x / den[:, None] - (c / b)[:, None] * y
    return li, gc[:, None] * dcdx, gc[:, None] * dcdy


def pair_score_loss(measure, loss_type, x, y, labels, margin=1.0, reduction="mean", dt=np.float64):
    """Whole step in explicit arithmetic: returns sim, probs, loss, dx, dy."""
    n = len(labels)
    scale = 1.0 / n if reduction == "mean" else 1.0
    s = score(measure, x, y, dt)
    p = probs(measure, s)
    if loss_type == "cosine":
        li, dx, dy = cosine_embedding(x, y, labels, margin, dt)
        dx, dy = dx * scale, dy * scale
    else:
        li, gi = loss_per_pair(loss_type, s, labels, margin)
        dx, dy = score_grad(measure, x, y, gi * scale, dt)
    loss = li.sum() * scale if reduction in ("mean", "sum") else li
    return s, p, loss, dx, dy


# ------------------------------------------------------------------ softmax ("cls") head + CE
def softmax_head(x, y, w, b, dt=np.float64):
    """a7: base.py:103-117 / numpy twin submit/similarity.py:19-24."""
    v = np.concatenate([_f(x, dt), _f(y, dt)], axis=1)
    logits = v @ _f(w, dt).T + _f(b, dt)
    m = logits.max(1, keepdims=True)
    e = np.exp(logits - m)
    return logits, e / e.sum(1, keepdims=True)


def softmax_head_ce(x, y, w, b, labels, dt=np.float64):
    """a7 + a12 (text.py:1408-1409,1473) with gradients: returns logits, probs, loss, dx, dy, dW, db."""
    x, y, w = _f(x, dt), _f(y, dt), _f(w, dt)
    n, h = x.shape
    logits, p = softmax_head(x, y, w, b, dt)
    labels = np.asarray(labels)
    loss = -np.log(p[np.arange(n), labels]).mean()
    dl = p.copy()
    dl[np.arange(n), labels] -= 1
    dl /= n
    v = np.concatenate([x, y], axis=1)
    dv = dl @ w
    return logits, p, loss, dv[:, :h], dv[:, h:], dl.T @ v, dl.sum(0)


# ------------------------------------------------------------------ retrieval keys / top-k
def orderable_u32(score_f32):
    """Monotone map fp32 -> u32 (larger float -> larger unsigned); the packing of the build's
    64-bit retrieval keys (SURVEY 7 "Hard parts" (c))."""
    u = np.asarray(score_f32, dtype=np.float32).view(np.uint32)
    return np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)


def pack_keys(score_f32, idx, descending=True):
    """key = (orderable(score) << 32) | (0xFFFFFFFF - idx) for 'larger is better' measures;
    for distances the score word is complemented so that ONE unsigned max-compare always
    means 'better score, then lower index'."""
    o = orderable_u32(score_f32).astype(np.uint64)
    if not descending:
        o = (~o) & np.uint64(0xFFFFFFFF)
    return (o << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.asarray(idx, dtype=np.uint64))


def unpack_keys(keys, descending=True):
    keys = np.asarray(keys, dtype=np.uint64)
    o = (keys >> np.uint64(32)).astype(np.uint32)
    if not descending:
        o = ~o
    u = np.where(o & 0x80000000, o & 0x7FFFFFFF, ~o).astype(np.uint32)
    idx = (np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF))).astype(np.int64)
    return u.view(np.float32), idx


def merge_topk_keys(parts, k):
    """Merge per-shard key lists [G, Q, k] -> [Q, k] by the unsigned compare (descending)."""
    g, q, kk = parts.shape
    allk = np.transpose(parts, (1, 0, 2)).reshape(q, g * kk)
    return np.sort(allk, axis=1)[:, ::-1][:, :k]


def score_matrix_inner(q, c, dt=np.float64):
    """All-pairs inner products (dense contraction form of a1)."""
    return _f(q, dt) @ _f(c, dt).T


def topk_stable(scores, k, descending=True):
    """Stable-sort top-k (ties -> lower index) on a score matrix."""
    s = np.asarray(scores)
    order = np.argsort(-s if descending else s, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(s, order, 1), order


def best_f1_and_threshold(scores, labels, high_score_more_similar=True):
    """finetune_bert.py:72-106 vectorised (numpy float64, the Python loop's operations in the loop's order): stable sort by
    score (np.argsort kind='stable' on the negated scores keeps equal scores in input order, like Python's sorted(...,
    reverse=True)), prefix counts, F1 of the cut points 0..n-2, FIRST maximum (the loop's strict '>').  Checked against
    the loop restatement torch_port.find_best_f1_and_threshold in tests/test_oracle_golden.py; exists so that 10^7 pairs
    can be checked in seconds."""
    s = np.asarray(scores, dtype=np.float64)
    lab = np.asarray(labels)
    n = len(s)
    if n < 2:
        return 0, 0, 0, 0, 0
    order = np.argsort(-s if high_score_more_similar else s, kind="stable")
    ss, ll = s[order], (lab[order] == 1)
    total = float(np.sum(lab))
    neg_total = n - total
    ncorrect = np.cumsum(ll)[: n - 1].astype(np.float64)
    nextract = np.arange(1, n, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        precision = ncorrect / nextract
        recall = ncorrect / total
        f1 = np.where(ncorrect > 0, 2 * precision * recall / (precision + recall), 0.0)
    f1 = np.nan_to_num(f1, nan=0.0)
    best = int(np.argmax(f1))
    if not f1[best] > 0:
        return 0, 0, 0, 0, 0
    fneg = nextract[best] - ncorrect[best]
    acc = (ncorrect[best] + neg_total - fneg) / n
    return float(acc), float(f1[best]), float(precision[best]), float(recall[best]), float((ss[best] + ss[best + 1]) / 2)


# ---------------------------------------------------------------------------------------------- dropout mask replay
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on numpy uint32 arrays (Salmon et al. 2011; the generator family of torch's CUDA dropout)."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0 = M0 * c0.astype(np.uint64)
        p1 = M1 * c2.astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint32(k0), lo1, hi0 ^ c3 ^ np.uint32(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_keep_mask(rows, cols, p, seed, step, stream):
    """The keep mask csrc/philox.cuh defines for a [rows, cols] tensor: element (row, col) = 16-bit lane (row & 7) of
    philox4x32_10(counter = (g_lo, g_hi, stream, step), key = seed), g = (row >> 3) * cols + col; kept <=> lane >= round(p * 2^16).
    Test infrastructure: replays the masks of ia_dropout_fwd / ia_project_tanh_fwd(dropout) on the host."""
    thr = min(65535, max(0, int(np.float32(p) * np.float32(65536.0) + np.float32(0.5))))
    groups = (rows + 7) // 8
    g = (np.arange(groups, dtype=np.uint64)[:, None] * np.uint64(cols) + np.arange(cols, dtype=np.uint64)[None, :])
    r = philox4x32_10((g & np.uint64(0xFFFFFFFF)).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
                      np.full(g.shape, stream, np.uint32), np.full(g.shape, step, np.uint32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    lanes = np.stack([r[0] & 0xFFFF, r[0] >> 16, r[1] & 0xFFFF, r[1] >> 16, r[2] & 0xFFFF, r[2] >> 16, r[3] & 0xFFFF, r[3] >> 16], axis=1)
    keep = lanes >= thr                                  # [groups, 8, cols]
    return keep.reshape(groups * 8, cols)[:rows]
