"""Functional API over the C ABI: torch tensors in, torch tensors out, autograd wired.

torch is plumbing here (device memory, streams, autograd graph); every score / loss / gradient value is
produced by the CUDA kernels of libia_b200.so.  Non-CUDA tensors raise: there is no CPU path.
"""
import torch

from . import _lib
from ._lib import LOSSES, MEASURES, REDUCTIONS, check, lib

_DT = {torch.float32: _lib.IA_F32, torch.bfloat16: _lib.IA_BF16, torch.float16: _lib.IA_F16}
_workspaces = {}


def _measure_id(measure):
    if measure not in MEASURES:
        raise ValueError(f"Unsupported similarty measure: {measure}")     # wording of reference base.py:64
    return MEASURES[measure]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _prep(x, y):
    if not (x.is_cuda and y.is_cuda):
        raise RuntimeError("item_alignment_b200 runs on CUDA tensors only (no CPU fallback)")
    if x.dim() != 2 or x.shape != y.shape:
        raise ValueError(f"expected two [N, D] tensors of equal shape, got {tuple(x.shape)} and {tuple(y.shape)}")
    if x.dtype != y.dtype:
        y = y.to(x.dtype)
    if x.dtype not in _DT:
        raise NotImplementedError(f"unsupported dtype {x.dtype}")
    if x.stride(1) != 1:
        x = x.contiguous()
    if y.stride(1) != 1:
        y = y.contiguous()
    return x, y


def workspace(device, nbytes=None):
    """Zero-initialised scratch for the deterministic loss reduction, one per (device, stream)."""
    need = nbytes or lib().ia_workspace_bytes()
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _ld(t):
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


# ------------------------------------------------------------------------------------------- raw ops
def pair_score_raw(measure, x, y, threshold=None, want_probs=True):
    """sim [, probs [, labels]] with no autograd (reference base.py:77-86, finetune_text.py:576-580)."""
    x, y = _prep(x, y)
    n, d = x.shape
    sim = torch.empty(n, dtype=torch.float32, device=x.device)
    probs = torch.empty(n, dtype=torch.float32, device=x.device) if (want_probs or threshold is not None) else None
    labels = torch.empty(n, dtype=torch.bool, device=x.device) if threshold is not None else None
    with torch.cuda.device(x.device):
        check(lib().ia_pair_score_fwd(_measure_id(measure), _DT[x.dtype], x.data_ptr(), y.data_ptr(), n, d, _ld(x), _ld(y),
                                      sim.data_ptr(), probs.data_ptr() if probs is not None else None,
                                      float(threshold) if threshold is not None else 0.0,
                                      labels.data_ptr() if labels is not None else None, _stream()))
    return sim, probs, labels


def pair_score_bwd_raw(measure, x, y, gsim, grad_dtype=None):
    x, y = _prep(x, y)
    n, d = x.shape
    gd = grad_dtype or x.dtype
    dx = torch.empty((n, d), dtype=gd, device=x.device)
    dy = torch.empty((n, d), dtype=gd, device=x.device)
    gsim = gsim.detach().to(torch.float32).contiguous()
    with torch.cuda.device(x.device):
        check(lib().ia_pair_score_bwd(_measure_id(measure), _DT[x.dtype], _DT[gd], x.data_ptr(), y.data_ptr(), _ld(x), _ld(y),
                                      gsim.data_ptr(), n, d, dx.data_ptr(), dy.data_ptr(), d, d, _stream()))
    return dx, dy


def _labels_i64(labels, n):
    labels = labels.view(-1)
    if labels.dtype != torch.int64:
        labels = labels.to(torch.int64)      # bce: float labels are accepted, {0,1} semantics
    labels = labels.contiguous()
    if labels.numel() != n:
        raise ValueError("labels must have one entry per pair")
    return labels


def pair_score_loss_raw(measure, loss_type, x, y, labels, margin=1.0, reduction="mean", grad_dtype=None,
                        grad_scale=1.0, want_grads=True):
    """ONE kernel: sim, probs, loss, dx, dy (reference head + ladder + backward; see include/ia_b200.h)."""
    x, y = _prep(x, y)
    n, d = x.shape
    if loss_type not in LOSSES:
        raise ValueError(f"unsupported loss_type for a vector-similarity head: {loss_type}")
    labels = _labels_i64(labels, n)
    gd = grad_dtype or x.dtype
    dev = x.device
    sim = torch.empty(n, dtype=torch.float32, device=dev)
    probs = torch.empty(n, dtype=torch.float32, device=dev)
    loss = torch.empty(n if reduction == "none" else 1, dtype=torch.float32, device=dev)
    dx = torch.empty((n, d), dtype=gd, device=dev) if want_grads else None
    dy = torch.empty((n, d), dtype=gd, device=dev) if want_grads else None
    with torch.cuda.device(dev):
        ws = workspace(dev)
        check(lib().ia_pair_score_loss_fwd_bwd(
            _measure_id(measure), LOSSES[loss_type], float(margin), REDUCTIONS[reduction], _DT[x.dtype], _DT[gd],
            x.data_ptr(), y.data_ptr(), _ld(x), _ld(y), labels.data_ptr(), n, d, sim.data_ptr(), probs.data_ptr(),
            loss.data_ptr(), dx.data_ptr() if want_grads else None, dy.data_ptr() if want_grads else None, d, d,
            float(grad_scale), None, 0, ws.data_ptr(), ws.numel(), _stream()))
    return sim, probs, (loss if reduction == "none" else loss[0]), dx, dy


def pair_score_loss_regrad_(measure, loss_type, x, y, labels, margin, reduction, upstream, dx=None, dy=None):
    """Gradients of the fused step for an upstream scalar that lives on the DEVICE (what autograd hands to backward():
    GradScaler's scale under --fp16, 1/accumulation_steps, ...).  The scalar is folded in before the single rounding to
    the gradient dtype -- 16-bit gradients neither underflow nor round twice (reference finetune_text.py:479-482 runs this
    backward in fp32 on the already scaled loss).  With dx, dy given (the buffers the forward launch filled for upstream
    == 1) the launch is a device-side no-op when the scalar is 1: no host sync, no traffic.  Returns (dx, dy)."""
    n, d = x.shape
    have = dx is not None
    if not have:
        dx = torch.empty((n, d), dtype=x.dtype, device=x.device)
        dy = torch.empty((n, d), dtype=x.dtype, device=x.device)
    up = upstream.detach().to(torch.float32).reshape(1).contiguous()
    with torch.cuda.device(x.device):
        ws = workspace(x.device)
        check(lib().ia_pair_score_loss_fwd_bwd(
            _measure_id(measure), LOSSES[loss_type], float(margin), REDUCTIONS[reduction], _DT[x.dtype], _DT[dx.dtype],
            x.data_ptr(), y.data_ptr(), _ld(x), _ld(y), labels.data_ptr(), n, d, None, None, None, dx.data_ptr(),
            dy.data_ptr(), d, d, 1.0, up.data_ptr(), int(have), ws.data_ptr(), ws.numel(), _stream()))
    return dx, dy


def scale_inplace_(a, b, g):
    """a *= g, b *= g with g a device scalar; the kernel is a no-op when g == 1 (no host sync)."""
    g = g.detach().to(torch.float32).reshape(1)
    with torch.cuda.device(a.device):
        check(lib().ia_scale_inplace(_DT[a.dtype], a.data_ptr(), b.data_ptr() if b is not None else None, a.numel(),
                                     g.data_ptr(), _stream()))


def score_loss_raw(loss_type, sim, target, margin=1.0, reduction="mean", want_grad=True):
    """HingeLoss / EuclideanDistanceLoss / BCEWithLogits on a score vector (reference loss.py)."""
    sim = sim.detach().to(torch.float32).contiguous().view(-1)
    target = target.view(-1).to(torch.int64).contiguous()
    n = sim.numel()
    out = torch.empty(n if reduction == "none" else 1, dtype=torch.float32, device=sim.device)
    g = torch.empty(n, dtype=torch.float32, device=sim.device) if want_grad else None
    with torch.cuda.device(sim.device):
        ws = workspace(sim.device)
        check(lib().ia_score_loss_fwd_bwd(LOSSES[loss_type], float(margin), REDUCTIONS[reduction], sim.data_ptr(),
                                          target.data_ptr(), n, out.data_ptr(), g.data_ptr() if want_grad else None, 1.0,
                                          ws.data_ptr(), ws.numel(), _stream()))
    return (out if reduction == "none" else out[0]), g


def softmax_head_raw(x, y, w, b, labels=None, grad_dtype=None, want_grads=True, grad_scale=1.0, upstream=None, grads=None,
                     want_wgrads=True):
    """TwoTowerClassificationHead (+ CrossEntropyLoss fwd/bwd when labels are given), reference base.py:103-117.
    upstream (device scalar) / grads (dx, dy, dw, db of an earlier launch): gradient recomputation of the autograd
    backward, see pair_score_loss_regrad_."""
    x, y = _prep(x, y)
    n, h = x.shape
    dev = x.device
    w = w.detach().to(torch.float32).contiguous()
    b = b.detach().to(torch.float32).contiguous()
    if w.shape != (2, 2 * h) or b.shape != (2,):
        raise NotImplementedError("the CUDA softmax head supports num_labels == 2")
    train = labels is not None
    gd = grad_dtype or x.dtype
    loss = dx = dy = dw = db = None
    ws_ptr, ws_n = None, 0
    regrad = upstream is not None
    up = upstream.detach().to(torch.float32).reshape(1).contiguous() if regrad else None
    with torch.cuda.device(dev):
        if regrad:
            if grads is not None:
                dx, dy, dw, db = grads
            else:
                dx = torch.empty((n, h), dtype=gd, device=dev)
                dy = torch.empty((n, h), dtype=gd, device=dev)
                dw = torch.empty((2, 2 * h), dtype=torch.float32, device=dev)
                db = torch.empty(2, dtype=torch.float32, device=dev)
            labels = labels.view(-1).to(torch.int64).contiguous()
            ws = workspace(dev, lib().ia_softmax_head_workspace_bytes(h))
            check(lib().ia_softmax_head_fwd_bwd(
                _DT[x.dtype], _DT[dx.dtype], x.data_ptr(), y.data_ptr(), _ld(x), _ld(y), w.data_ptr(), b.data_ptr(),
                labels.data_ptr(), n, h, None, None, None, dx.data_ptr(), dy.data_ptr(), h, h, dw.data_ptr(), db.data_ptr(),
                1.0, up.data_ptr(), int(grads is not None), ws.data_ptr(), ws.numel(), _stream()))
            return None, None, None, dx, dy, dw, db
        logits = torch.empty((n, 2), dtype=torch.float32, device=dev)
        probs = torch.empty((n, 2), dtype=torch.float32, device=dev)
        if train:
            labels = labels.view(-1).to(torch.int64).contiguous()
            loss = torch.empty(1, dtype=torch.float32, device=dev)
            if want_grads:
                dx = torch.empty((n, h), dtype=gd, device=dev)
                dy = torch.empty((n, h), dtype=gd, device=dev)
                if want_wgrads:      # False: frozen head (dx, dy only; the dW / db finalize launch is skipped)
                    dw = torch.empty((2, 2 * h), dtype=torch.float32, device=dev)
                    db = torch.empty(2, dtype=torch.float32, device=dev)
            ws = workspace(dev, lib().ia_softmax_head_workspace_bytes(h))
            ws_ptr, ws_n = ws.data_ptr(), ws.numel()
        check(lib().ia_softmax_head_fwd_bwd(
            _DT[x.dtype], _DT[gd], x.data_ptr(), y.data_ptr(), _ld(x), _ld(y), w.data_ptr(), b.data_ptr(),
            labels.data_ptr() if train else None, n, h, logits.data_ptr(), probs.data_ptr(),
            loss.data_ptr() if train else None, dx.data_ptr() if dx is not None else None,
            dy.data_ptr() if dy is not None else None, h, h, dw.data_ptr() if dw is not None else None,
            db.data_ptr() if db is not None else None, float(grad_scale), None, 0, ws_ptr, ws_n, _stream()))
    return logits, probs, (loss[0] if train else None), dx, dy, dw, db


# ------------------------------------------------------------------------------------------- gather-and-score
def _prep_gather(emb_x, emb_y, src_idx, tgt_idx, check_indices=True):
    if not (emb_x.is_cuda and emb_y.is_cuda):
        raise RuntimeError("item_alignment_b200 runs on CUDA tensors only (no CPU fallback)")
    if emb_x.dim() != 2 or emb_y.dim() != 2 or emb_x.shape[1] != emb_y.shape[1] or emb_x.dtype != emb_y.dtype:
        raise ValueError("expected two [M, D] embedding matrices of equal D and dtype")
    if emb_x.dtype not in _DT:
        raise NotImplementedError(f"unsupported dtype {emb_x.dtype}")
    if emb_x.stride(1) != 1:
        emb_x = emb_x.contiguous()
    if emb_y.stride(1) != 1:
        emb_y = emb_y.contiguous()
    src_idx = src_idx.to(device=emb_x.device, dtype=torch.int64).contiguous().view(-1)
    tgt_idx = tgt_idx.to(device=emb_x.device, dtype=torch.int64).contiguous().view(-1)
    if src_idx.numel() != tgt_idx.numel():
        raise ValueError("src_idx and tgt_idx must have the same length")
    if check_indices and src_idx.numel() > 0:
        # what torch indexing in the reference loop (graph.py:87-117) would raise; one small reduction + one host read.
        # (The kernels are memory-safe without it: an out-of-range index poisons that pair with NaN.)
        bad = ((src_idx < 0) | (src_idx >= emb_x.shape[0]) | (tgt_idx < 0) | (tgt_idx >= emb_y.shape[0])).any()
        if bool(bad):
            raise IndexError(f"pair index out of range for embedding matrices with {emb_x.shape[0]} / {emb_y.shape[0]} rows")
    return emb_x, emb_y, src_idx, tgt_idx


def pair_score_gather_raw(measure, emb_x, emb_y, src_idx, tgt_idx, threshold=None, check_indices=True):
    """Scores of pairs given as row indices into embedding matrices, one launch (replaces the per-pair loop of
    reference src/models/graph.py:87-117).  Returns sim, probs, labels-or-None.  Out-of-range indices raise IndexError
    (check_indices=False skips the host-side check: such pairs then come back as NaN)."""
    emb_x, emb_y, src_idx, tgt_idx = _prep_gather(emb_x, emb_y, src_idx, tgt_idx, check_indices)
    n, d = src_idx.numel(), emb_x.shape[1]
    dev = emb_x.device
    sim = torch.empty(n, dtype=torch.float32, device=dev)
    probs = torch.empty(n, dtype=torch.float32, device=dev)
    labels = torch.empty(n, dtype=torch.bool, device=dev) if threshold is not None else None
    with torch.cuda.device(dev):
        check(lib().ia_pair_score_gather_fwd(_measure_id(measure), _DT[emb_x.dtype], emb_x.data_ptr(), emb_y.data_ptr(),
                                             emb_x.shape[0], emb_y.shape[0], _ld(emb_x), _ld(emb_y), src_idx.data_ptr(),
                                             tgt_idx.data_ptr(), n, d, sim.data_ptr(), probs.data_ptr(),
                                             float(threshold) if threshold is not None else 0.0,
                                             labels.data_ptr() if labels is not None else None, _stream()))
    return sim, probs, labels


def pair_score_loss_gather_raw(measure, loss_type, emb_x, emb_y, src_idx, tgt_idx, labels, margin=1.0, reduction="mean",
                               grad_dtype=None, want_grads=True, check_indices=True):
    """Fused gather + score + loss + backward; dx, dy are dense per pair ([n, D])."""
    emb_x, emb_y, src_idx, tgt_idx = _prep_gather(emb_x, emb_y, src_idx, tgt_idx, check_indices)
    n, d = src_idx.numel(), emb_x.shape[1]
    dev = emb_x.device
    if loss_type not in LOSSES:
        raise ValueError(f"unsupported loss_type for a vector-similarity head: {loss_type}")
    labels = labels.view(-1).to(torch.int64).contiguous()
    gd = grad_dtype or emb_x.dtype
    sim = torch.empty(n, dtype=torch.float32, device=dev)
    probs = torch.empty(n, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    dx = torch.empty((n, d), dtype=gd, device=dev) if want_grads else None
    dy = torch.empty((n, d), dtype=gd, device=dev) if want_grads else None
    with torch.cuda.device(dev):
        ws = workspace(dev)
        check(lib().ia_pair_score_gather_loss_fwd_bwd(
            _measure_id(measure), LOSSES[loss_type], float(margin), REDUCTIONS[reduction], _DT[emb_x.dtype], _DT[gd],
            emb_x.data_ptr(), emb_y.data_ptr(), emb_x.shape[0], emb_y.shape[0], _ld(emb_x), _ld(emb_y), src_idx.data_ptr(),
            tgt_idx.data_ptr(), labels.data_ptr(), n, d, sim.data_ptr(), probs.data_ptr(), loss.data_ptr(),
            dx.data_ptr() if want_grads else None, dy.data_ptr() if want_grads else None, d, d, 1.0, ws.data_ptr(),
            ws.numel(), _stream()))
    return sim, probs, loss[0], dx, dy


class _GatherPairLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb_x, emb_y, src_idx, tgt_idx, labels, measure, loss_type, margin, reduction):
        need = emb_x.requires_grad or emb_y.requires_grad
        # per-pair gradients in fp32: they are scattered (index_add) in fp32 anyway, and the upstream scalar (GradScaler,
        # 1/accum) is applied out of place to the fp32 sums before the one rounding to the embedding dtype
        sim, probs, loss, dx, dy = pair_score_loss_gather_raw(measure, loss_type, emb_x, emb_y, src_idx, tgt_idx, labels, margin,
                                                              reduction, grad_dtype=torch.float32, want_grads=need)
        if need:
            ctx.save_for_backward(dx, dy, src_idx, tgt_idx)
        ctx.shapes = (emb_x.shape, emb_y.shape, emb_x.dtype, emb_y.dtype)
        ctx.mark_non_differentiable(sim, probs)
        return loss, sim, probs

    @staticmethod
    def backward(ctx, gloss, _gs, _gp):
        dx, dy, src_idx, tgt_idx = ctx.saved_tensors
        sx, sy, tx, ty = ctx.shapes
        g = gloss.detach().to(torch.float32)
        gx = (torch.zeros(sx, dtype=torch.float32, device=dx.device).index_add_(0, src_idx, dx) * g).to(tx)
        gy = (torch.zeros(sy, dtype=torch.float32, device=dy.device).index_add_(0, tgt_idx, dy) * g).to(ty)
        return gx, gy, None, None, None, None, None, None, None


def pair_score_loss_gather(measure, loss_type, emb_x, emb_y, src_idx, tgt_idx, labels, margin=1.0, reduction="mean"):
    """(sim, probs, loss) for index pairs; loss.backward() scatters the per-pair gradients into the embedding
    matrices' gradients (index_add)."""
    loss, sim, probs = _GatherPairLossFn.apply(emb_x, emb_y, src_idx.to(emb_x.device).long().view(-1),
                                               tgt_idx.to(emb_x.device).long().view(-1), labels, measure, loss_type,
                                               float(margin), reduction)
    return sim, probs, loss


# ------------------------------------------------------------------------------------------- evaluation sweep
def threshold_sweep_counts(probs, labels, thresholds):
    """[T, 4] int64 confusion counts (tp, fp, fn, tn) of `probs >= thresholds[k]` (reference finetune_text.py:576-580)."""
    if not probs.is_cuda:
        raise RuntimeError("item_alignment_b200 runs on CUDA tensors only (no CPU fallback)")
    probs = probs.detach().to(torch.float32).contiguous().view(-1)
    labels = labels.to(device=probs.device, dtype=torch.int64).contiguous().view(-1)
    thr = torch.as_tensor([float(t) for t in thresholds], dtype=torch.float64, device=probs.device)
    if thr.numel() < 1 or thr.numel() > 32:
        raise ValueError("1..32 thresholds per sweep")
    counts = torch.empty((thr.numel(), 4), dtype=torch.int64, device=probs.device)
    with torch.cuda.device(probs.device):
        check(lib().ia_threshold_sweep(probs.data_ptr(), labels.data_ptr(), probs.numel(), thr.data_ptr(), thr.numel(),
                                       counts.data_ptr(), _stream()))
    return counts


def row_inv_norm(x, eps=1e-8):
    if x.stride(1) != 1:
        x = x.contiguous()
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().ia_row_inv_norm(_DT[x.dtype], x.data_ptr(), x.shape[0], x.shape[1], _ld(x), float(eps), out.data_ptr(), _stream()))
    return out


def _prep_project(f1, f2, w, b):
    if not (f1.is_cuda and f2.is_cuda and w.is_cuda):
        raise RuntimeError("item_alignment_b200 runs on CUDA tensors only (no CPU fallback)")
    if f1.dim() != 2 or f1.shape != f2.shape or w.dim() != 2 or w.shape[1] != f1.shape[1]:
        raise ValueError(f"expected features [N, K] x2 and weight [H, K], got {tuple(f1.shape)}, {tuple(f2.shape)}, {tuple(w.shape)}")
    if f1.dtype not in (torch.bfloat16, torch.float16):
        raise NotImplementedError("the fused projection runs bf16 / fp16 features on the tensor cores; keep torch for fp32")
    f2 = f2.to(f1.dtype)
    w = w.detach().to(f1.dtype)            # nn.Linear under autocast casts its fp32 master weight the same way
    f1, f2, w = (t if (t.stride(1) == 1 and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0) else t.contiguous() for t in (f1, f2, w))
    if b is not None:
        b = b.detach().to(torch.float32).contiguous()
    return f1, f2, w, b


def project_tanh_raw(f1, f2, w, b, fast_tanh=False):
    """x = tanh(f1 @ w.T + b), y = tanh(f2 @ w.T + b) from ONE tcgen05 GEMM launch (reference base.py:67-75 with
    dropout inactive); outputs in the feature dtype.  fast_tanh=True selects the hardware tanh.approx (2^-11 relative, opt-in)."""
    f1, f2, w, b = _prep_project(f1, f2, w, b)
    n, k = f1.shape
    h = w.shape[0]
    x = torch.empty((n, h), dtype=f1.dtype, device=f1.device)
    y = torch.empty((n, h), dtype=f1.dtype, device=f1.device)
    with torch.cuda.device(f1.device):
        check(lib().ia_project_tanh_fwd(_DT[f1.dtype], f1.data_ptr(), f2.data_ptr(), _ld(f1), _ld(f2), n, k, w.data_ptr(), _ld(w),
                                        b.data_ptr() if b is not None else None, h, x.data_ptr(), y.data_ptr(), h, h, int(bool(fast_tanh)),
                                        _stream()))
    return x, y


def project_score_raw(measure, f1, f2, w, b, threshold=None, want_embeds=False, fast_tanh=False):
    """Projection + pair score in one launch (+ a [N]-sized finalize): (sim, probs, labels, x, y).  With
    want_embeds=False the embeddings never reach HBM."""
    f1, f2, w, b = _prep_project(f1, f2, w, b)
    n, k = f1.shape
    h = w.shape[0]
    dev = f1.device
    x = torch.empty((n, h), dtype=f1.dtype, device=dev) if want_embeds else None
    y = torch.empty((n, h), dtype=f1.dtype, device=dev) if want_embeds else None
    sim = torch.empty(n, dtype=torch.float32, device=dev)
    probs = torch.empty(n, dtype=torch.float32, device=dev)
    labels = torch.empty(n, dtype=torch.bool, device=dev) if threshold is not None else None
    mid = _measure_id(measure)
    if measure == "softmax":
        raise ValueError(f"Unsupported similarty measure: {measure}")
    with torch.cuda.device(dev):
        need = lib().ia_project_score_workspace_bytes(n, k, h)
        ws = torch.empty(max(int(need), 16), dtype=torch.uint8, device=dev)
        check(lib().ia_project_score_fwd(mid, _DT[f1.dtype], f1.data_ptr(), f2.data_ptr(), _ld(f1), _ld(f2), n, k, w.data_ptr(),
                                         _ld(w), b.data_ptr() if b is not None else None, h,
                                         x.data_ptr() if want_embeds else None, y.data_ptr() if want_embeds else None, h, h,
                                         sim.data_ptr(), probs.data_ptr(), float(threshold) if threshold is not None else 0.0,
                                         labels.data_ptr() if labels is not None else None, int(bool(fast_tanh)), ws.data_ptr(), ws.numel(),
                                         _stream()))
    return sim, probs, labels, x, y


# ------------------------------------------------------------------------------------------- projection, training side
def _keep_scale(p):
    return 1.0 / (1.0 - float(p)) if p > 0 else 1.0


def dropout_raw(x, p, seed, step, stream_id):
    """out = x * mask / (1 - p) with the counter-based masks of csrc/philox.cuh (stream_id 0: f1, 1: f2)."""
    if x.dtype not in (torch.bfloat16, torch.float16) or x.dim() != 2 or x.shape[1] % 8 != 0:
        raise NotImplementedError("the dropout kernel runs [rows, cols] bf16 / fp16 tensors with cols % 8 == 0")
    x = x if (x.stride(1) == 1 and x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0) else x.contiguous()
    out = torch.empty_like(x, memory_format=torch.contiguous_format)
    with torch.cuda.device(x.device):
        check(lib().ia_dropout_fwd(_DT[x.dtype], x.data_ptr(), _ld(x), x.shape[0], x.shape[1], float(p), int(seed), int(step), int(stream_id),
                                   out.data_ptr(), out.shape[1], _stream()))
    return out


def project_tanh_dropout_raw(f1d, f2d, w, b, p, seed, step):
    """x, y = drop(tanh(f @ w.T + b)) for both sides from ONE tcgen05 GEMM launch, output dropout in the epilogue (streams 2, 3).
    f1d, f2d are the (already dropped) features."""
    f1d, f2d, w, b = _prep_project(f1d, f2d, w, b)
    n, k = f1d.shape
    h = w.shape[0]
    x = torch.empty((n, h), dtype=f1d.dtype, device=f1d.device)
    y = torch.empty((n, h), dtype=f1d.dtype, device=f1d.device)
    with torch.cuda.device(f1d.device):
        check(lib().ia_project_tanh_dropout_fwd(_DT[f1d.dtype], f1d.data_ptr(), f2d.data_ptr(), _ld(f1d), _ld(f2d), n, k, w.data_ptr(), _ld(w),
                                                b.data_ptr() if b is not None else None, h, x.data_ptr(), y.data_ptr(), h, h, float(p),
                                                int(seed), int(step), _stream()))
    return x, y


def tanh_dropout_bwd_raw(g, out, p):
    """d_pre = g * keep / (1 - p) * (1 - tanh^2) from the head's output `out` = drop(tanh(.)) and an upstream gradient g."""
    g = g.to(out.dtype).contiguous()
    dpre = torch.empty_like(out, memory_format=torch.contiguous_format)
    with torch.cuda.device(out.device):
        check(lib().ia_tanh_dropout_bwd(_DT[out.dtype], g.data_ptr(), g.shape[1], out.data_ptr(), _ld(out), out.shape[0], out.shape[1],
                                        _keep_scale(p), dpre.data_ptr(), dpre.shape[1], _stream()))
    return dpre


def transpose16_raw(w):
    """w [h, k] -> w.T materialised [k, h] (16-bit); the data-gradient GEMM wants the contraction dimension contiguous."""
    w = w.contiguous()
    wt = torch.empty((w.shape[1], w.shape[0]), dtype=w.dtype, device=w.device)
    with torch.cuda.device(w.device):
        check(lib().ia_transpose16(w.data_ptr(), w.shape[0], w.shape[1], w.shape[1], wt.data_ptr(), wt.shape[1], _stream()))
    return wt


def project_dgrad_raw(d1, d2, wt, p, seed, step):
    """df1, df2 = (d_pre @ w) * input-dropout mask / (1 - p): the forward GEMM kernel on (d_pre, w.T) with the mask in its epilogue."""
    n, h = d1.shape
    k = wt.shape[0]
    df1 = torch.empty((n, k), dtype=d1.dtype, device=d1.device)
    df2 = torch.empty((n, k), dtype=d1.dtype, device=d1.device)
    with torch.cuda.device(d1.device):
        check(lib().ia_project_dgrad(_DT[d1.dtype], d1.data_ptr(), d2.data_ptr(), _ld(d1), _ld(d2), n, h, wt.data_ptr(), _ld(wt), k,
                                     df1.data_ptr(), df2.data_ptr(), k, k, float(p), int(seed), int(step), _stream()))
    return df1, df2


def project_wgrad_raw(d1, d2, f1d, f2d, want_db=True):
    """dw [h, k] fp32 = d1.T @ f1d + d2.T @ f2d, db [h] fp32 = column sums of d1, d2: one tcgen05 GEMM contracting over the rows."""
    n, h = d1.shape
    k = f1d.shape[1]
    dev = d1.device
    d1, d2, f1d, f2d = (t if (t.stride(1) == 1 and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0) else t.contiguous() for t in (d1, d2, f1d, f2d))
    if _ld(d1) != _ld(d2):
        d2 = d2.contiguous(); d1 = d1.contiguous()
    if _ld(f1d) != _ld(f2d):
        f1d = f1d.contiguous(); f2d = f2d.contiguous()
    dw = torch.empty((h, k), dtype=torch.float32, device=dev)
    db = torch.empty(h, dtype=torch.float32, device=dev) if want_db else None
    with torch.cuda.device(dev):
        need = lib().ia_project_wgrad_workspace_bytes(n, h, k)
        ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 256
        check(lib().ia_project_wgrad(_DT[d1.dtype], d1.data_ptr(), d2.data_ptr(), _ld(d1), f1d.data_ptr(), f2d.data_ptr(), _ld(f1d), n, h, k,
                                     dw.data_ptr(), db.data_ptr() if want_db else None, ws.data_ptr() + off, need, _stream()))
    return dw, db


def _project_backward(ctx_f1d, ctx_f2d, w16, d1, d2, p, seed, step, needs):
    """Shared tail of the projection backward: data gradient (dgrad) and weight / bias gradients (wgrad) on the tensor cores."""
    df1 = df2 = dw = db = None
    if needs[0] or needs[1]:
        df1, df2 = project_dgrad_raw(d1, d2, transpose16_raw(w16), p, seed, step)
    if needs[2] or needs[3]:
        dw, db = project_wgrad_raw(d1, d2, ctx_f1d, ctx_f2d, want_db=needs[3])
    return df1, df2, dw, db


class _ProjectTrainFn(torch.autograd.Function):
    """(x, y) = drop(tanh(dense(drop(f1)))), drop(tanh(dense(drop(f2)))) -- VecSimClassificationHead.forward in train() mode
    (reference base.py:50-53,67-75) -- with every GEMM of forward AND backward on tcgen05 kernels: forward GEMM with bias + tanh +
    output dropout in its epilogue, data gradient through the same kernel on (d_pre, W^T) with the input mask in its epilogue,
    weight gradient through the MN-major split-K GEMM.  Dropout masks are counter-based (seed, step): nothing is stored."""

    @staticmethod
    def forward(ctx, f1, f2, w, b, p, seed, step):
        w16 = w.detach().to(f1.dtype)
        f1d = dropout_raw(f1, p, seed, step, 0) if p > 0 else f1
        f2d = dropout_raw(f2, p, seed, step, 1) if p > 0 else f2
        x, y = project_tanh_dropout_raw(f1d, f2d, w16, b, p, seed, step)
        ctx.save_for_backward(f1d, f2d, w16, x, y)
        ctx.cfg = (float(p), int(seed), int(step), w.dtype, b.dtype if b is not None else None)
        return x, y

    @staticmethod
    def backward(ctx, gx, gy):
        f1d, f2d, w16, x, y = ctx.saved_tensors
        p, seed, step, wdt, bdt = ctx.cfg
        d1 = tanh_dropout_bwd_raw(gx if gx is not None else torch.zeros_like(x), x, p)
        d2 = tanh_dropout_bwd_raw(gy if gy is not None else torch.zeros_like(y), y, p)
        needs = (ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2], bdt is not None and ctx.needs_input_grad[3])
        df1, df2, dw, db = _project_backward(f1d, f2d, w16, d1, d2, p, seed, step, needs)
        return (df1 if needs[0] else None, df2 if needs[1] else None, dw.to(wdt) if needs[2] else None,
                db.to(bdt) if needs[3] else None, None, None, None)


class _ProjectScoreLossTrainFn(torch.autograd.Function):
    """Head projection (train mode) + pair score + loss ladder + their whole backward: after the forward GEMM the fused pair-loss
    launch writes d_pre (the gradient at the dense layer's output: loss backward, similarity backward, dropout and tanh backward in
    one HBM pass, ia_pair_score_loss_act_bwd); backward() only runs the two gradient GEMMs.  An upstream scalar != 1 (GradScaler,
    gradient accumulation) re-issues the pair launch with the scalar folded in before the rounding, like _FusedPairLossFn."""

    @staticmethod
    def forward(ctx, f1, f2, w, b, labels, measure, loss_type, margin, p, seed, step):
        w16 = w.detach().to(f1.dtype)
        f1d = dropout_raw(f1, p, seed, step, 0) if p > 0 else f1
        f2d = dropout_raw(f2, p, seed, step, 1) if p > 0 else f2
        x, y = project_tanh_dropout_raw(f1d, f2d, w16, b, p, seed, step)
        n, h = x.shape
        labels = _labels_i64(labels, n)
        sim = torch.empty(n, dtype=torch.float32, device=x.device)
        probs = torch.empty(n, dtype=torch.float32, device=x.device)
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        d1 = torch.empty_like(x)
        d2 = torch.empty_like(y)
        with torch.cuda.device(x.device):
            ws = workspace(x.device)
            check(lib().ia_pair_score_loss_act_bwd(
                _measure_id(measure), LOSSES[loss_type], float(margin), REDUCTIONS["mean"], _DT[x.dtype], _DT[x.dtype], x.data_ptr(), y.data_ptr(),
                h, h, labels.data_ptr(), n, h, sim.data_ptr(), probs.data_ptr(), loss.data_ptr(), d1.data_ptr(), d2.data_ptr(), h, h, 1.0, None, 0,
                _keep_scale(p), ws.data_ptr(), ws.numel(), _stream()))
        ctx.save_for_backward(f1d, f2d, w16, x, y, labels)
        ctx.first = (d1, d2)
        ctx.cfg = (float(p), int(seed), int(step), w.dtype, b.dtype if b is not None else None, measure, loss_type, float(margin))
        ctx.mark_non_differentiable(sim, probs)
        ctx.set_materialize_grads(False)
        return x, y, sim, probs, loss[0]

    @staticmethod
    def backward(ctx, gx, gy, _gs, _gp, gloss):
        f1d, f2d, w16, x, y, labels = ctx.saved_tensors
        p, seed, step, wdt, bdt, measure, loss_type, margin = ctx.cfg
        n, h = x.shape
        first, ctx.first = ctx.first, None
        have = first is not None
        d1, d2 = first if have else (torch.empty_like(x), torch.empty_like(y))
        if gloss is None:
            d1.zero_(); d2.zero_()
        else:
            up = gloss.detach().to(torch.float32).reshape(1).contiguous()
            with torch.cuda.device(x.device):
                ws = workspace(x.device)
                check(lib().ia_pair_score_loss_act_bwd(
                    _measure_id(measure), LOSSES[loss_type], margin, REDUCTIONS["mean"], _DT[x.dtype], _DT[x.dtype], x.data_ptr(), y.data_ptr(),
                    h, h, labels.data_ptr(), n, h, None, None, None, d1.data_ptr(), d2.data_ptr(), h, h, 1.0, up.data_ptr(), int(have),
                    _keep_scale(p), ws.data_ptr(), ws.numel(), _stream()))
        if gx is not None:        # the embeddings were also used elsewhere: add that path's contribution
            d1 = d1 + tanh_dropout_bwd_raw(gx, x, p)
        if gy is not None:
            d2 = d2 + tanh_dropout_bwd_raw(gy, y, p)
        needs = (ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2], bdt is not None and ctx.needs_input_grad[3])
        df1, df2, dw, db = _project_backward(f1d, f2d, w16, d1, d2, p, seed, step, needs)
        return (df1 if needs[0] else None, df2 if needs[1] else None, dw.to(wdt) if needs[2] else None,
                db.to(bdt) if needs[3] else None, None, None, None, None, None, None, None)


def project_tanh_train(f1, f2, w, b, p, seed, step):
    """Training-mode projection of both sides (dropout p on the features and on the tanh outputs), differentiable."""
    f1, f2, _, _ = _prep_project(f1, f2, w, b)
    return _ProjectTrainFn.apply(f1, f2, w, b, float(p), int(seed), int(step))


def project_score_loss_train(measure, loss_type, f1, f2, w, b, labels, p, seed, step, margin=1.0):
    """(x, y, sim, probs, loss) of the training step from the encoder features: projection with dropout, pair score, loss ladder;
    loss.backward() runs the fused gradient launch's d_pre through the two gradient GEMMs."""
    if loss_type not in LOSSES:
        raise ValueError(f"unsupported loss_type for a vector-similarity head: {loss_type}")
    f1, f2, _, _ = _prep_project(f1, f2, w, b)
    return _ProjectScoreLossTrainFn.apply(f1, f2, w, b, labels, measure, loss_type, float(margin), float(p), int(seed), int(step))


# ------------------------------------------------------------------------------------------- autograd
class _PairScoreFn(torch.autograd.Function):
    """sim = similarity(x, y) with the backward kernel (unfused head -> loss module sequence of the reference)."""

    @staticmethod
    def forward(ctx, x, y, measure):
        sim, _, _ = pair_score_raw(measure, x, y, want_probs=False)
        ctx.save_for_backward(x, y)
        ctx.measure = measure
        return sim

    @staticmethod
    def backward(ctx, gsim):
        x, y = ctx.saved_tensors
        dx, dy = pair_score_bwd_raw(ctx.measure, x, y, gsim)
        return dx.to(x.dtype), dy.to(y.dtype), None


class _FusedPairLossFn(torch.autograd.Function):
    """(loss, sim, probs) from ONE kernel that also produced dx, dy for an upstream gradient of 1.

    backward(): the common case (a plain loss.backward()) hands those buffers over untouched -- the gradient launch it
    issues is a device-side no-op when the upstream scalar is 1, and there is no host sync.  Any other upstream scalar
    (GradScaler under --fp16, loss / accumulation_steps) recomputes dx, dy from x, y with the scalar folded in BEFORE the
    single rounding to the input dtype, which is what the reference's fp32 backward on the scaled loss does
    (finetune_text.py:479-482); same traffic as rescaling the buffers would cost, but fp16 gradients of a mean over 64k
    pairs do not flush to zero first.  The forward's buffers are handed out once; a second backward through the same node
    (retain_graph=True) recomputes into fresh buffers, so nothing is ever scaled twice."""

    @staticmethod
    def forward(ctx, x, y, labels, measure, loss_type, margin, reduction):
        need = x.requires_grad or y.requires_grad
        labels = _labels_i64(labels, x.shape[0])
        sim, probs, loss, dx, dy = pair_score_loss_raw(measure, loss_type, x, y, labels, margin, reduction, want_grads=need)
        if need:
            ctx.save_for_backward(x, y, labels)
            ctx.first = (dx, dy)
            ctx.cfg = (measure, loss_type, margin, reduction)
        ctx.mark_non_differentiable(sim, probs)
        return loss, sim, probs

    @staticmethod
    def backward(ctx, gloss, _gsim, _gprobs):
        x, y, labels = ctx.saved_tensors
        first, ctx.first = ctx.first, None
        dx, dy = first if first is not None else (None, None)
        dx, dy = pair_score_loss_regrad_(*ctx.cfg[:2], x, y, labels, ctx.cfg[2], ctx.cfg[3], gloss, dx, dy)
        return dx, dy, None, None, None, None, None


class _ScoreLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sim, target, loss_type, margin, reduction):
        out, g = score_loss_raw(loss_type, sim, target, margin, reduction, want_grad=sim.requires_grad)
        if g is not None:
            ctx.save_for_backward(g)
        ctx.reduction = reduction
        ctx.n = sim.numel()
        ctx.in_dtype = sim.dtype
        return out

    @staticmethod
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        return (g * gout).to(ctx.in_dtype), None, None, None, None


class _SoftmaxHeadLogitsFn(torch.autograd.Function):
    """logits of the two-tower softmax head; arbitrary upstream d/dlogits handled with library GEMMs
    (this is the unfused drop-in path; the fused CE path is softmax_head_ce)."""

    @staticmethod
    def forward(ctx, x, y, w, b):
        logits, _, _, _, _, _, _ = softmax_head_raw(x, y, w, b)
        ctx.save_for_backward(x, y, w)
        return logits

    @staticmethod
    def backward(ctx, gl):
        x, y, w = ctx.saved_tensors
        h = x.shape[1]
        gl = gl.float()
        wf = w.float()
        dx = (gl @ wf[:, :h]).to(x.dtype)
        dy = (gl @ wf[:, h:]).to(y.dtype)
        dw = torch.cat((gl.t() @ x.float(), gl.t() @ y.float()), dim=1).to(w.dtype)
        return dx, dy, dw, gl.sum(0)


class _FusedSoftmaxCEFn(torch.autograd.Function):
    """Same contract as _FusedPairLossFn: the forward launch's gradients are handed out once, untouched when the upstream
    scalar is 1 (device-side no-op); any other scalar, or a second backward, recomputes with the scalar folded in."""

    @staticmethod
    def forward(ctx, x, y, w, b, labels):
        labels = labels.view(-1).to(torch.int64).contiguous()
        logits, probs, loss, dx, dy, dw, db = softmax_head_raw(x, y, w, b, labels)
        ctx.save_for_backward(x, y, w, b, labels)
        ctx.first = (dx, dy, dw, db)
        ctx.wdtype, ctx.bdtype = w.dtype, b.dtype
        ctx.mark_non_differentiable(logits, probs)
        return loss, logits, probs

    @staticmethod
    def backward(ctx, gloss, _gl, _gp):
        x, y, w, b, labels = ctx.saved_tensors
        first, ctx.first = ctx.first, None
        _, _, _, dx, dy, dw, db = softmax_head_raw(x, y, w, b, labels, upstream=gloss, grads=first)
        return dx, dy, dw.to(ctx.wdtype), db.to(ctx.bdtype), None


# ------------------------------------------------------------------------------------------- public
def pair_similarity(measure, x, y):
    """similarity(x, y) -> [N] fp32, differentiable (reference base.py:54-62,77)."""
    if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad):
        x, y = _prep(x, y)
        return _PairScoreFn.apply(x, y, measure)
    return pair_score_raw(measure, x, y, want_probs=False)[0]


def probs_of(measure, sim):
    """reference base.py:79-86 on an autograd-tracked sim (tiny [N] torch ops)."""
    if measure == "cosine":
        return (sim + 1) / 2
    if measure in ("l1", "l2"):
        return torch.exp(-sim)
    if measure == "inner_product":
        return torch.sigmoid(sim)
    raise ValueError(f"Unsupported similarty measure: {measure}")


def pair_score(measure, x, y, threshold=None):
    """(sim, probs[, labels]): one kernel when no gradient is needed."""
    if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad):
        sim = pair_similarity(measure, x, y)
        probs = probs_of(measure, sim)
        if threshold is None:
            return sim, probs
        return sim, probs, probs.detach().double() >= float(threshold)
    sim, probs, labels = pair_score_raw(measure, x, y, threshold)
    return (sim, probs) if threshold is None else (sim, probs, labels)


def pair_score_loss(measure, loss_type, x, y, labels, margin=1.0, reduction="mean"):
    """Fused head score + loss ladder + backward: returns (sim, probs, loss); loss.backward() hands the
    gradients computed in the same pass to autograd."""
    if reduction == "none":
        sim = pair_similarity(measure, x, y)
        probs = probs_of(measure, sim)
        if loss_type == "cosine":
            raise NotImplementedError("reduction='none' is not offered for the cosine-embedding loss")
        target = labels if loss_type == "bce" else labels * 2 - 1
        return sim, probs, _ScoreLossFn.apply(sim, target, loss_type, margin, "none")
    x, y = _prep(x, y)
    loss, sim, probs = _FusedPairLossFn.apply(x, y, labels, measure, loss_type, float(margin), reduction)
    return sim, probs, loss


def score_loss(loss_type, sim, target, margin=1.0, reduction="mean"):
    return _ScoreLossFn.apply(sim, target, loss_type, float(margin), reduction)


def softmax_head(x, y, w, b):
    """(logits, probs) of the two-tower softmax head, differentiable."""
    if torch.is_grad_enabled() and any(t.requires_grad for t in (x, y, w, b)):
        x, y = _prep(x, y)
        logits = _SoftmaxHeadLogitsFn.apply(x, y, w, b)
        return logits, torch.softmax(logits, dim=1)
    logits, probs, _, _, _, _, _ = softmax_head_raw(x, y, w, b)
    return logits, probs


def softmax_head_ce(x, y, w, b, labels):
    """Fused softmax head + CrossEntropyLoss forward/backward: (logits, probs, loss)."""
    x, y = _prep(x, y)
    loss, logits, probs = _FusedSoftmaxCEFn.apply(x, y, w, b, labels)
    return logits, probs, loss


def project_tanh(f1, f2, w, b):
    """(x, y) = tanh(linear(f1, w, b)), tanh(linear(f2, w, b)), differentiable (reference base.py:67-75, eval mode)."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (f1, f2, w, b)):
        return project_tanh_train(f1, f2, w, b, 0.0, 0, 0)       # p = 0: same forward kernel, backward on the gradient GEMMs
    return project_tanh_raw(f1, f2, w, b)


def project_score(measure, f1, f2, w, b, threshold=None, want_embeds=False, fast_tanh=False):
    """Inference: projection + similarity + probability map (+ threshold labels) from the encoder features in one
    launch.  Returns (sim, probs[, labels]) or, with want_embeds, (x, y, sim, probs[, labels])."""
    sim, probs, labels, x, y = project_score_raw(measure, f1, f2, w, b, threshold, want_embeds, fast_tanh)
    out = (sim, probs) if threshold is None else (sim, probs, labels)
    return ((x, y) + out) if want_embeds else out
