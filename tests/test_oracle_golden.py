"""CPU: both oracle layers against the golden vectors produced by the real reference modules
(tests/golden/make_golden.py) and against the reference's three weak artefacts."""
import numpy as np
import pytest
import torch

from oracle import formula, ref_import, torch_port
from tests import parity

MEASURES = ("inner_product", "cosine", "l1", "l2")
TAGS = ("a", "b", "c", "d", "e", "f")


def _combos():
    for m in MEASURES:
        for l in ("bce", "hinge", "euclidean", "cosine"):
            for margin in ((1.0, 0.3) if l in ("hinge", "cosine") else (1.0,)):
                yield m, l, margin


def _close(a, b, rtol, atol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin) or np.allclose(a[fin], b[fin], rtol=rtol, atol=atol)
    np.testing.assert_allclose(a[fin], b[fin], rtol=rtol, atol=atol)


@pytest.mark.parametrize("tag", TAGS)
def test_torch_port_matches_reference_golden(golden, tag):
    g = golden("pair_golden")
    x, y = torch.from_numpy(g[f"{tag}/x"]), torch.from_numpy(g[f"{tag}/y"])
    labels = torch.from_numpy(g[f"{tag}/labels"])
    for m, l, margin in _combos():
        key = f"{tag}/{m}/{l}/{margin}"
        sim, probs, loss, dx, dy = torch_port.pair_score_loss_fwd_bwd(m, l, x, y, labels, margin)
        # same torch ops in the same order as the reference modules -> bit-identical on one thread,
        # allow a few ulp for thread-count dependent reduction order
        _close(sim, g[key + "/sim"], 2e-6, 1e-6)
        _close(probs, g[key + "/probs"], 2e-6, 1e-7)
        _close(loss, g[key + "/loss"], 1e-5, 1e-7)
        _close(dx, g[key + "/dx"], 1e-5, 1e-6 * max(1.0, float(np.nanmax(np.abs(g[key + "/dx"][np.isfinite(g[key + "/dx"])]), initial=0))))
        _close(dy, g[key + "/dy"], 1e-5, 1e-6 * max(1.0, float(np.nanmax(np.abs(g[key + "/dy"][np.isfinite(g[key + "/dy"])]), initial=0))))


@pytest.mark.parametrize("tag", TAGS)
def test_formula_matches_reference_golden(golden, tag):
    """Explicit float64 formulas vs the reference's fp32 outputs, under the fp32 tolerance rules of
    tests/parity.py (the same rules the CUDA kernels are held to)."""
    g = golden("pair_golden")
    x, y, labels = g[f"{tag}/x"], g[f"{tag}/y"], g[f"{tag}/labels"]
    n = len(labels)
    rtol = 1e-5
    for m, l, margin in _combos():
        key = f"{tag}/{m}/{l}/{margin}"
        s, p, loss, dx, dy = formula.pair_score_loss(m, l, x, y, labels, margin)
        parity.assert_scores_close(m, s, g[key + "/sim"], x, y, rtol)
        parity.assert_probs_close(p, g[key + "/probs"], rtol, parity.score_atol(m, x, y, rtol))
        # euclidean on a near-zero score (1/s, -1/s^2) amplifies the fp32 rounding of s without bound
        tiny = (np.abs(s) < 1e-2) if l == "euclidean" else np.zeros(n, bool)
        if not tiny.any():
            parity.assert_loss_close(loss, g[key + "/loss"], 10 * rtol,
                                     atol=10 * float(parity.score_atol(m, x, y, rtol).max()))
        dsdx, dsdy = parity.grad_term_scale(m if l != "cosine" else "cosine", x, y)
        parity.assert_grad_close(dx, g[key + "/dx"], dsdx, 1.0 / n, 10 * rtol, key + "/dx", skip_rows=tiny)
        parity.assert_grad_close(dy, g[key + "/dy"], dsdy, 1.0 / n, 10 * rtol, key + "/dy", skip_rows=tiny)


def test_edge_rows_of_fixture_a(golden):
    """zero rows -> cosine 0; identical rows -> PairwiseDistance gives eps*sqrt(D) / eps*D, not 0."""
    g = golden("pair_golden")
    d = g["a/x"].shape[1]
    cos = g["a/cosine/bce/1.0/sim"]
    assert cos[0] == 0 and cos[3] == 0
    np.testing.assert_allclose(g["a/l2/bce/1.0/sim"][1], 1e-6 * np.sqrt(d), rtol=1e-4)
    np.testing.assert_allclose(g["a/l1/bce/1.0/sim"][1], 1e-6 * d, rtol=1e-4)
    s = formula.score("cosine", g["a/x"], g["a/y"])
    assert s[0] == 0 and s[3] == 0


def test_heads_golden(golden):
    g = golden("head_golden")
    for m in MEASURES:
        k = f"vecsim/{m}/"
        x, y, sim, probs = torch_port.vecsim_head(m, torch.from_numpy(g[k + "f1"]), torch.from_numpy(g[k + "f2"]),
                                                  torch.from_numpy(g[k + "w"]), torch.from_numpy(g[k + "b"]))
        _close(x, g[k + "x"], 1e-6, 1e-7)
        _close(sim, g[k + "sim"], 2e-6, 1e-6)
        _close(probs, g[k + "probs"], 2e-6, 1e-7)
    for tag in ("s", "m"):
        k = f"twotower/{tag}/"
        t = {n: torch.from_numpy(g[k + n]) for n in ("f1", "f2", "w", "b", "labels")}
        logits, probs, loss, dx, dy, dw, db = torch_port.softmax_head_ce_fwd_bwd(t["f1"], t["f2"], t["w"], t["b"], t["labels"])
        for ours, name in ((logits, "logits"), (probs, "probs"), (loss, "loss"), (dx, "dx"), (dy, "dy"), (dw, "dw"), (db, "db")):
            _close(ours, g[k + name], 1e-5, 1e-7)
        f = formula.softmax_head_ce(g[k + "f1"], g[k + "f2"], g[k + "w"], g[k + "b"], g[k + "labels"])
        for ours, name in zip(f, ("logits", "probs", "loss", "dx", "dy", "dw", "db")):
            _close(ours, g[k + name], 2e-5, 2e-7)


def test_retrieval_golden(golden):
    g = golden("retrieval_golden")
    q, c, k = torch.from_numpy(g["q"]), torch.from_numpy(g["c"]), int(g["k"])
    for m in MEASURES:
        sc = torch_port.all_pairs_scores(m, q, c).numpy()
        if m in ("inner_product", "l1"):
            # exact-arithmetic fixture: sums of multiples of 1/2 (+eps for l1 is NOT exact) ...
            pass
        np.testing.assert_allclose(sc, g[f"{m}/scores"], rtol=2e-6, atol=2e-6)
        if m == "inner_product":
            assert np.array_equal(sc, g[f"{m}/scores"])          # exact sums: order independent
            v, i = torch_port.retrieve_topk(m, q, c, k)
            assert np.array_equal(i.numpy(), g[f"{m}/top_idx"])  # bit-exact incl. ties -> lower index
            assert np.array_equal(v.numpy(), g[f"{m}/top_scores"])
            fv, fi = formula.topk_stable(formula.score_matrix_inner(g["q"], g["c"]), k)
            assert np.array_equal(fi, g[f"{m}/top_idx"])


def test_key_packing_orders_like_stable_sort(golden):
    g = golden("retrieval_golden")
    k = int(g["k"])
    for m in MEASURES:
        desc = m in ("inner_product", "cosine")
        sc = g[f"{m}/scores"]
        idx = np.broadcast_to(np.arange(sc.shape[1]), sc.shape)
        keys = formula.pack_keys(sc, idx, descending=desc)
        top = np.sort(keys, axis=1)[:, ::-1][:, :k]
        s, i = formula.unpack_keys(top, descending=desc)
        assert np.array_equal(i, g[f"{m}/top_idx"])
        assert np.array_equal(s, g[f"{m}/top_scores"])
        # merge of shard-wise top-k == global top-k
        parts = []
        for lo in range(0, sc.shape[1], 100):
            kk = formula.pack_keys(sc[:, lo:lo + 100], idx[:, lo:lo + 100], descending=desc)
            parts.append(np.sort(kk, axis=1)[:, ::-1][:, :k])
        merged = formula.merge_topk_keys(np.stack(parts), k)
        assert np.array_equal(merged, top)


def test_submit_artefacts(golden):
    g = golden("submit_golden")
    tgt0, thr = g["deepai/tgt0"], g["deepai/threshold"]
    assert len(tgt0) == 15909
    labels = np.array([torch_port.compute_passthrough([0.0], [float(v)]) >= t for v, t in zip(tgt0, thr)])
    assert labels.sum() == 5319                       # known answer
    assert np.array_equal(labels, g["deepai/labels"])
    for i in range(4):
        w, b = g[f"softmax/{i}/w"], g[f"softmax/{i}/b"]
        for e1, e2, p1 in zip(g[f"softmax/{i}/e1"], g[f"softmax/{i}/e2"], g[f"softmax/{i}/p1"]):
            assert torch_port.compute_softmax_head(list(e1), list(e2), w, b) == p1
            _, p = formula.softmax_head(e1[None], e2[None], w, b)
            np.testing.assert_allclose(p[0, 1], p1, rtol=1e-12)
            _, pt = torch_port.two_tower_head(torch.from_numpy(e1[None]), torch.from_numpy(e2[None]),
                                              torch.from_numpy(w), torch.from_numpy(b))
            np.testing.assert_allclose(pt[0, 1].item(), p1, rtol=2e-5)


def test_pure_python_inner_matches():
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(768), rng.standard_normal(768)
    s = torch_port.compute_inner(list(a), list(b), bias=0.25)
    np.testing.assert_allclose(s, float(a @ b) + 0.25, rtol=1e-12)
    np.testing.assert_allclose(formula.score("inner_product", a[None], b[None])[0], s - 0.25, rtol=1e-12)


def test_threshold_label_rule():
    """labels = probs >= thr on fp32 probs compared in float64 (finetune_text.py:576-580); for cosine
    a tiny negative sim still labels positive at 0.5 because fp32 (sim+1)/2 rounds to 0.5."""
    sim = torch.tensor([-1e-9, -1e-6, 0.0, 1e-9], dtype=torch.float32)
    probs = torch_port.probs_of("cosine", sim)
    assert torch_port.threshold_labels(probs, 0.5).tolist() == [True, False, True, True]
    thr = np.arange(0.1, 1.0, 0.1)[2]                # 0.30000000000000004
    p = torch.tensor([np.float32(0.3), np.nextafter(np.float32(0.3), np.float32(1))])
    assert torch_port.threshold_labels(p, thr).tolist() == [bool(np.float64(np.float32(0.3)) >= thr), True]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_live_reference_cross_check():
    """When /root/reference is present: oracle vs the live reference modules on fresh random inputs."""
    import types
    base, rl = ref_import.base(), ref_import.loss()
    gen = torch.Generator().manual_seed(7)
    x, y = torch.tanh(torch.randn(64, 96, generator=gen)), torch.tanh(torch.randn(64, 96, generator=gen))
    t = (torch.rand(64, generator=gen) < 0.5).long() * 2 - 1
    for m in MEASURES:
        cfg = types.SimpleNamespace(cls_layers="1", cls_pool="cls", hidden_size=96, classifier_dropout=0.0,
                                    hidden_dropout_prob=0.0, similarity_measure=m)
        head = base.VecSimClassificationHead(cfg)
        sim = head.similarity(x, y)
        assert torch.equal(sim, torch_port.similarity(m, x, y))
        assert torch.equal(rl.HingeLoss(0.7)(sim, t), torch_port.hinge_loss(sim, t, 0.7))
        assert torch.equal(rl.EuclideanDistanceLoss()(sim, t), torch_port.euclidean_loss(sim, t))


def test_eval_restatements_small_cases():
    """finetune_bert.py:72-106 and finetune_text.py:576-580 restatements on hand-checkable inputs."""
    scores = [0.9, 0.8, 0.7, 0.6, 0.5, 0.4]
    labels = [1, 1, 0, 1, 0, 0]
    acc, f1, p, r, thr = torch_port.find_best_f1_and_threshold(scores, labels)
    assert (p, r) == (0.75, 1.0) and abs(f1 - 6 / 7) < 1e-15 and abs(thr - 0.55) < 1e-15 and abs(acc - 5 / 6) < 1e-15
    pr, rc, f = torch_port.threshold_sweep(np.array([0.2, 0.6, 0.7, 0.1], dtype=np.float32), np.array([0, 1, 0, 1]), [0.5])
    assert (pr[0], rc[0], f[0]) == (0.5, 0.5, 0.5)
    sims, probs, loss = torch_port.gcn_pair_loop("cosine", torch.eye(3), [dict(src_idx=0, tgt_idx=0, item_label=1),
                                                                         dict(src_idx=0, tgt_idx=1, item_label=0)], "hinge", 1.0)
    assert sims.tolist() == [1.0, 0.0] and probs.tolist() == [1.0, 0.5] and float(loss) == 0.5


def test_vectorised_best_f1_equals_the_loop_restatement():
    """oracle/formula.best_f1_and_threshold (used to check 10^7 pairs on the GPU) == the loop restatement of
    finetune_bert.py:72-106, ties and both sort directions included."""
    from oracle import formula, torch_port
    rng = np.random.default_rng(3)
    for n, ties, dt in ((2, False, np.float64), (3, True, np.float32), (1000, False, np.float32), (2503, True, np.float64),
                        (4000, True, np.float32)):
        labels = (rng.random(n) < 0.35).astype(np.int64)
        scores = (rng.standard_normal(n) + labels).astype(dt)
        if ties:
            scores = (np.round(scores * 3) / 3).astype(dt)
        for high in (True, False):
            ref = torch_port.find_best_f1_and_threshold([float(v) for v in scores], labels.tolist(), high)
            got = formula.best_f1_and_threshold(scores, labels, high)
            assert tuple(float(v) for v in ref) == tuple(float(v) for v in got), (n, ties, high)
    assert formula.best_f1_and_threshold(np.zeros(5), np.zeros(5, dtype=np.int64)) == (0, 0, 0, 0, 0)     # no positives
