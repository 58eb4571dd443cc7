"""GPU parity of the rows SURVEY 8(f) marks "next": gather-and-score for index pairs (the GCN per-pair loop,
reference src/models/graph.py:87-117), the threshold sweep of finetune_text.py:576-580 and the best-F1 threshold
search of finetune_bert.py:72-106."""
import numpy as np
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MEASURES = ("inner_product", "cosine", "l1", "l2")


@pytest.mark.parametrize("dt,d", [(torch.float32, 96), (torch.bfloat16, 1024), (torch.float32, 50)])
def test_gather_and_score_vs_per_pair_loop(dt, d):
    import item_alignment_b200.functional as F_
    from oracle import torch_port
    gen = torch.Generator().manual_seed(11 + d)
    m_nodes, n = 301, 513
    emb = torch.tanh(torch.randn(m_nodes, d, generator=gen)).to(dt)
    src = torch.randint(0, m_nodes, (n,), generator=gen)
    tgt = torch.randint(0, m_nodes, (n,), generator=gen)
    tgt[:7] = src[:7]                                      # self pairs
    labels = (torch.rand(n, generator=gen) < 0.5).long()
    pairs = [dict(src_idx=int(s), tgt_idx=int(t), item_label=int(l)) for s, t, l in zip(src, tgt, labels)]
    embd = emb.to(DEV)
    rtol = 1e-5 if dt == torch.float32 else 1e-3
    for m in MEASURES:
        rs, rp, rl = torch_port.gcn_pair_loop(m, emb, pairs, "hinge", 0.5)
        sim, probs, lab = F_.pair_score_gather_raw(m, embd, embd, src, tgt, threshold=0.5)
        x, y = emb[src].float(), emb[tgt].float()
        parity.assert_scores_close(m, sim, rs, x, y, rtol)
        parity.assert_probs_close(probs, rp, rtol, parity.score_atol(m, x, y, rtol))
        assert np.array_equal(lab.cpu().numpy(), torch_port.threshold_labels(probs, 0.5))
        # identical to the dense kernels on materialised rows, bit for bit
        s2, p2, _ = F_.pair_score_raw(m, embd[src.to(DEV)], embd[tgt.to(DEV)])
        assert torch.equal(sim, s2) and torch.equal(probs, p2)
        # fused loss + gradients scattered into the embedding matrix
        e = embd.clone().float().requires_grad_(True) if dt == torch.float32 else embd.clone().requires_grad_(True)
        _, _, loss = F_.pair_score_loss_gather(m, "hinge", e, e, src, tgt, labels.to(DEV), margin=0.5)
        loss.backward()
        parity.assert_loss_close(loss, rl, 10 * rtol, 10 * float(parity.score_atol(m, x, y, rtol).max()))
        er = emb.float().clone().requires_grad_(True)
        sim_r = torch_port.similarity(m, er[src], er[tgt])
        torch_port.loss_ladder("hinge", sim_r, er[src], er[tgt], labels, 0.5).backward()
        scale = float(er.grad.abs().max())
        tol = (2e-4 if dt == torch.float32 else 2e-2) * max(scale, 1e-6)
        assert float((e.grad.float().cpu() - er.grad).abs().max()) <= tol, m


def test_threshold_sweep_vs_sklearn_loop():
    import item_alignment_b200 as ia
    from oracle import torch_port
    gen = torch.Generator().manual_seed(3)
    n = 100_003
    labels = (torch.rand(n, generator=gen) < 0.3).long()
    probs = torch.sigmoid(torch.randn(n, generator=gen) + 1.5 * (labels.float() - 0.5))
    probs[:50] = torch.tensor(np.float32(0.3))                     # values sitting exactly on a float32-rounded threshold
    probs[50:100] = torch.tensor(np.nextafter(np.float32(0.3), np.float32(1)))
    thresholds = np.arange(0.1, 1.0, 0.1)
    p, r, f = ia.threshold_sweep(probs.to(DEV), labels.to(DEV), thresholds)
    rp, rr, rf = torch_port.threshold_sweep(probs.numpy(), labels.numpy(), thresholds)
    np.testing.assert_allclose(p, rp, rtol=1e-13); np.testing.assert_allclose(r, rr, rtol=1e-13); np.testing.assert_allclose(f, rf, rtol=1e-13)
    # the counts themselves are exact integers
    c = ia.functional.threshold_sweep_counts(probs.to(DEV), labels.to(DEV), thresholds).cpu().numpy()
    for k, t in enumerate(thresholds):
        pred = probs.numpy() >= t
        lab = labels.numpy().astype(bool)
        assert tuple(c[k]) == (int((pred & lab).sum()), int((pred & ~lab).sum()), int((~pred & lab).sum()), int((~pred & ~lab).sum()))
    # degenerate: no predicted positives -> sklearn's 0.0 convention
    p, r, f = ia.threshold_sweep(torch.zeros(10, device=DEV), torch.ones(10, dtype=torch.long, device=DEV), [0.5])
    assert (p[0], r[0], f[0]) == (0.0, 0.0, 0.0)


def test_find_best_f1_and_threshold_vs_reference_loop():
    import item_alignment_b200 as ia
    from oracle import torch_port
    gen = torch.Generator().manual_seed(5)
    for n, ties in ((2000, False), (3001, True), (2, False)):
        labels = (torch.rand(n, generator=gen) < 0.4).long()
        scores = torch.randn(n, generator=gen, dtype=torch.float64) + labels.double()
        if ties:
            scores = torch.round(scores * 4) / 4                # many equal scores: stable order must be kept
        for high in (True, False):
            ours = ia.find_best_f1_and_threshold(scores.to(DEV), labels.to(DEV), high)
            ref = torch_port.find_best_f1_and_threshold(scores.tolist(), labels.tolist(), high)
            assert ours == tuple(float(v) for v in ref), (n, ties, high, ours, ref)


def test_gather_rejects_out_of_range_indices():
    """a negative or >= M item id must neither read out of bounds nor pass silently: the Python mirror raises
    IndexError (what torch indexing in the reference loop does); straight through the C ABI the pair is poisoned with NaN."""
    import item_alignment_b200.functional as F_
    emb = torch.tanh(torch.randn(40, 64, device=DEV))
    src = torch.tensor([0, 5, 39, 7], device=DEV)
    tgt = torch.tensor([1, 40, 2, -1], device=DEV)          # 40 and -1 are out of range
    labels = torch.tensor([1, 0, 1, 0], device=DEV)
    with pytest.raises(IndexError):
        F_.pair_score_gather_raw("cosine", emb, emb, src, tgt)
    with pytest.raises(IndexError):
        F_.pair_score_loss_gather("cosine", "bce", emb, emb, src, tgt, labels)
    for m in MEASURES:
        sim, probs, _ = F_.pair_score_gather_raw(m, emb, emb, src, tgt, check_indices=False)
        assert torch.isnan(sim[[1, 3]]).all() and torch.isfinite(sim[[0, 2]]).all()
        ref = F_.pair_score_raw(m, emb[src[[0, 2]]], emb[tgt[[0, 2]]])[0]
        assert torch.equal(sim[[0, 2]], ref)
        out = F_.pair_score_loss_gather_raw(m, "hinge", emb, emb, src, tgt, labels, check_indices=False)
        assert torch.isnan(out[2]) and torch.isfinite(out[3][[0, 2]]).all()


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_best_f1_kernels_ties_tiles_and_ten_million_pairs(dt):
    """ia_best_f1_threshold (radix sort + scan + F1 + arg-max kernels) against the oracle: tile boundaries (8192 per CTA),
    heavy ties (stable order must be kept), both directions, degenerate inputs, and 10^7 pairs."""
    import item_alignment_b200 as ia
    from oracle import formula, torch_port
    gen = torch.Generator().manual_seed(17)
    for n, ties in ((2, False), (3, True), (255, True), (8192, False), (8193, True), (24577, True), (100_003, False)):
        labels = (torch.rand(n, generator=gen) < 0.4).long()
        scores = (torch.randn(n, generator=gen, dtype=torch.float64) + labels.double()).to(dt)
        if ties:
            scores = (torch.round(scores * 2) / 2).to(dt)          # a handful of distinct values
        for high in (True, False):
            ours = ia.find_best_f1_and_threshold(scores.to(DEV), labels.to(DEV), high)
            ref = formula.best_f1_and_threshold(scores.numpy(), labels.numpy(), high)
            assert tuple(float(v) for v in ours) == tuple(float(v) for v in ref), (n, ties, high, ours, ref)
            if n <= 8193:
                loop = torch_port.find_best_f1_and_threshold(scores.tolist(), labels.tolist(), high)
                assert tuple(float(v) for v in ours) == tuple(float(v) for v in loop)
    # all-negative labels: no cut has F1 > 0
    assert ia.find_best_f1_and_threshold(torch.randn(5000, device=DEV).to(dt), torch.zeros(5000, dtype=torch.long, device=DEV)) == (0, 0, 0, 0, 0)
    # 10^7 pairs with ties (bf16-rounded probabilities: ~ 10^4 distinct values)
    n = 10_000_000
    g = torch.Generator(device=DEV).manual_seed(23)
    labels = (torch.rand(n, device=DEV, generator=g) < 0.2).long()
    scores = torch.sigmoid(torch.randn(n, device=DEV, generator=g) + 2.0 * labels.float()).to(torch.bfloat16).to(dt)
    for high in (True, False):
        ours = ia.find_best_f1_and_threshold(scores if high else -scores, labels, high)
        ref = formula.best_f1_and_threshold((scores if high else -scores).cpu().numpy(), labels.cpu().numpy(), high)
        assert tuple(float(v) for v in ours) == tuple(float(v) for v in ref), (high, ours, ref)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(3):
        ia.find_best_f1_and_threshold(scores, labels, True)
    print(f"best-F1 search over 10^7 {dt} scores: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per call (incl. result read-back)")
