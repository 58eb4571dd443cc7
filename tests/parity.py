"""Tolerance rules shared by the CPU (oracle) and GPU (CUDA vs oracle) parity tests.

north_star: scores / losses / gradients within 1e-3 relative for bf16 inputs and 1e-5 for fp32,
labels and top-k indices bit-exact.  "Relative" is made well-posed as in SURVEY 7 "Hard parts":

* inner / cosine scores cancel, so |a-b| <= rtol*|b| + rtol*|x||y| (inner) or + rtol (cosine in [-1,1]);
  l1 / l2 are cancellation-free (+ rtol * 1e-6*D for the eps floor).
* gradients: dL/ds = sigma(s)-l (bce) is itself a cancelling fp32 difference, so every gradient row
  is compared with |a-b| <= rtol*|b| + rtol*max(rowmax|b|, scale*rowmax|ds/dx|).
"""
import numpy as np

RTOL = {"fp32": 1e-5, "bf16": 1e-3, "fp16": 1e-3}


def _np(a):
    if hasattr(a, "detach"):
        a = a.detach().float().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


def score_atol(measure, x, y, rtol):
    x, y = _np(x), _np(y)
    if measure == "inner_product":
        return rtol * np.linalg.norm(x, axis=1) * np.linalg.norm(y, axis=1) + 1e-30
    if measure == "cosine":
        return np.full(x.shape[0], rtol)
    return np.full(x.shape[0], rtol * 1e-6 * x.shape[1])


def assert_scores_close(measure, ours, ref, x, y, rtol, what="sim"):
    ours, ref = _np(ours), _np(ref)
    atol = score_atol(measure, x, y, rtol)
    err = np.abs(ours - ref)
    bound = rtol * np.abs(ref) + atol
    bad = ~(err <= bound) & ~(np.isnan(ours) & np.isnan(ref)) & ~((ours == ref))
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} outside tolerance; worst err {np.nanmax(err / bound):.3g}x bound"


def assert_probs_close(ours, ref, rtol, sim_atol=None):
    """probs = f(sim): error bound = rtol*|p| + |f'| * (sim error) <= rtol*|p| + sim_atol (|f'| <= 1)."""
    ours, ref = _np(ours), _np(ref)
    extra = 0.0 if sim_atol is None else sim_atol
    bad = ~(np.abs(ours - ref) <= rtol * np.abs(ref) + extra + 1e-37) & ~(ours == ref)
    assert not bad.any(), f"probs: {bad.sum()} of {bad.size} outside tolerance"


def assert_loss_close(ours, ref, rtol, atol=0.0):
    ours, ref = float(_np(ours)), float(_np(ref))
    if np.isnan(ref) or np.isinf(ref):
        assert (np.isnan(ours) and np.isnan(ref)) or ours == ref, (ours, ref)
        return
    assert abs(ours - ref) <= rtol * abs(ref) + atol, f"loss {ours} vs {ref} (rtol {rtol}, atol {atol})"


def grad_term_scale(measure, x, y):
    """Per-row magnitude of the UNCANCELLED terms of d sim/d x (the conditioning scale of the row):
    inner: max|y|; cosine / cosine-embedding: max|y^|/max(|x|,eps) (yh - s*xh cancels when x ~ y);
    l1 / l2: 1.  Returns (scale_for_dx, scale_for_dy), each [N,1]."""
    x, y = _np(x), _np(y)
    if measure == "inner_product":
        return np.abs(y).max(1, keepdims=True), np.abs(x).max(1, keepdims=True)
    if measure == "cosine":
        nx = np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-8)
        ny = np.maximum(np.linalg.norm(y, axis=1, keepdims=True), 1e-8)
        ax = np.maximum(np.abs(x).max(1, keepdims=True) / nx, 1e-30)
        ay = np.maximum(np.abs(y).max(1, keepdims=True) / ny, 1e-30)
        return np.maximum(ax, ay) / nx, np.maximum(ax, ay) / ny
    one = np.ones((x.shape[0], 1))
    return one, one


def assert_grad_close(ours, ref, term_scale, scale, rtol, what="grad", skip_rows=None):
    """ours/ref [N,D]; term_scale [N,1] from grad_term_scale; scale = 1/N for a mean loss (|dL/ds| <= 1
    for bce/hinge; larger upstream magnitudes show up in rowmax|ref| instead)."""
    ours, ref = _np(ours), _np(ref)
    fin = np.isfinite(ref)
    if skip_rows is not None:
        fin &= ~np.asarray(skip_rows)[:, None]
    rowmax = np.max(np.where(fin, np.abs(ref), 0.0), axis=1, keepdims=True)
    dmax = _np(term_scale) * scale
    bound = rtol * np.abs(ref) + rtol * np.maximum(rowmax, dmax) + 1e-37
    err = np.abs(ours - ref)
    bad = fin & ~(err <= bound)
    assert not bad.any(), (f"{what}: {bad.sum()} of {bad.size} outside tolerance; worst "
                           f"{np.nanmax(np.where(fin, err / bound, 0)):.3g}x bound at row {np.argwhere(bad)[0][0]}")


def assert_topk_matches(scores, idx, ref_matrix, k, descending, tol, exact=False):
    """Top-k parity against a full oracle score matrix [Q, C].

    exact=True: indices and scores bit-identical to the stable-sort top-k of ref_matrix.
    Otherwise (GPU accumulation order != CPU order): per query
      * scores match the oracle's sorted scores within tol,
      * every returned row's oracle score is within tol of qualifying, every oracle row that beats the k-th
        by more than tol is returned,
      * wherever the oracle's neighbouring ranks are separated by more than 2*tol the index matches exactly.
    Returns the number of (query, rank) positions that were gap-ambiguous."""
    scores, idx = _np(scores), np.asarray(idx.detach().cpu().numpy() if hasattr(idx, "detach") else idx)
    ref = np.asarray(ref_matrix, dtype=np.float64)
    q, c = ref.shape
    kk = min(k, c)
    order = np.argsort(-ref if descending else ref, axis=1, kind="stable")[:, :kk]
    top = np.take_along_axis(ref, order, 1)
    if exact:
        assert np.array_equal(idx[:, :kk], order), "top-k indices differ from the stable-sort oracle"
        assert np.array_equal(scores[:, :kk].astype(np.float32), top.astype(np.float32)), "top-k scores differ"
        return 0
    assert np.all(np.abs(scores[:, :kk] - top) <= tol + 1e-6 * np.abs(top)), \
        f"sorted scores differ by up to {np.abs(scores[:, :kk] - top).max():.3g} (tol {tol})"
    sgn = -1.0 if descending else 1.0
    ambiguous = 0
    for i in range(q):
        kth = top[i, kk - 1]
        got = idx[i, :kk]
        assert len(set(got.tolist())) == kk, f"query {i}: duplicate rows in the result"
        got_ref = ref[i, got]
        assert np.all(sgn * (got_ref - kth) <= tol + 1e-12), f"query {i}: a returned row does not qualify"
        must = np.nonzero(sgn * (ref[i] - kth) < -tol)[0]
        assert np.isin(must, got).all(), f"query {i}: a clearly better row is missing"
        gaps_prev = np.abs(np.diff(top[i], prepend=top[i, 0] - sgn * 1e9))
        gaps_next = np.abs(np.diff(top[i], append=(ref[i, np.argsort(sgn * ref[i], kind='stable')[kk]] if c > kk else top[i, -1] + sgn * 1e9)))
        clear = (gaps_prev > 2 * tol) & (gaps_next > 2 * tol)
        assert np.array_equal(got[clear], order[i][clear]), f"query {i}: index mismatch at a gap-separated rank"
        ambiguous += int((~clear).sum())
    return ambiguous
