"""GPU parity of the pair-scoring kernels (through the ctypes C-ABI) against the golden vectors produced
by the real reference modules, against the CPU oracle on seeded inputs, and -- at BASELINE's full sizes --
through size-independent properties."""
import numpy as np
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu

MEASURES = ("inner_product", "cosine", "l1", "l2")
LOSSES = ("bce", "hinge", "euclidean", "cosine")
DEV = "cuda:0"
TAG_DTYPE = {"a": torch.float32, "b": torch.float32, "c": torch.bfloat16, "d": torch.float16,
             "e": torch.float32, "f": torch.float32}


@pytest.fixture(scope="module")
def F_():
    import item_alignment_b200.functional as f
    return f


def _rtol(dt):
    return 1e-5 if dt == torch.float32 else 1e-3


def _margins(l):
    return (1.0, 0.3) if l in ("hinge", "cosine") else (1.0,)


@pytest.mark.parametrize("tag", list(TAG_DTYPE))
def test_fused_fwd_bwd_vs_reference_golden(golden, F_, tag):
    """Every measure x loss of the ladder on the fixtures the reference modules produced.  Gradients are
    requested in fp32 (tolerance check) and in the input dtype (must equal the fp32 result rounded once)."""
    g = golden("pair_golden")
    dt = TAG_DTYPE[tag]
    rtol = _rtol(dt)
    x = torch.from_numpy(g[f"{tag}/x"]).to(DEV).to(dt)
    y = torch.from_numpy(g[f"{tag}/y"]).to(DEV).to(dt)
    labels = torch.from_numpy(g[f"{tag}/labels"]).to(DEV)
    xr, yr = g[f"{tag}/x"], g[f"{tag}/y"]
    n = len(labels)
    for m in MEASURES:
        for l in LOSSES:
            for margin in _margins(l):
                key = f"{tag}/{m}/{l}/{margin}"
                sim, probs, loss, dx, dy = F_.pair_score_loss_raw(m, l, x, y, labels, margin, "mean", grad_dtype=torch.float32)
                parity.assert_scores_close(m, sim, g[key + "/sim"], xr, yr, rtol, key + "/sim")
                parity.assert_probs_close(probs, g[key + "/probs"], rtol, parity.score_atol(m, xr, yr, rtol))
                s64 = g[key + "/sim"].astype(np.float64)
                tiny = (np.abs(s64) < 1e-2) if l == "euclidean" else np.zeros(n, bool)
                if not tiny.any():
                    parity.assert_loss_close(loss, g[key + "/loss"], 10 * rtol, 10 * float(parity.score_atol(m, xr, yr, rtol).max()))
                tx, ty = parity.grad_term_scale(m if l != "cosine" else "cosine", xr, yr)
                parity.assert_grad_close(dx, g[key + "/dx"], tx, 1.0 / n, 10 * rtol, key + "/dx", skip_rows=tiny)
                parity.assert_grad_close(dy, g[key + "/dy"], ty, 1.0 / n, 10 * rtol, key + "/dy", skip_rows=tiny)
                if dt != torch.float32:
                    _, _, loss2, dxl, dyl = F_.pair_score_loss_raw(m, l, x, y, labels, margin, "mean")
                    assert dxl.dtype == dt and torch.equal(loss2, loss)
                    assert torch.equal(dxl, dx.to(dt)) and torch.equal(dyl, dy.to(dt))   # same fp32 value, rounded once


@pytest.mark.parametrize("tag", ["a", "c", "f"])
def test_forward_and_unfused_backward_vs_golden(golden, F_, tag):
    g = golden("pair_golden")
    dt = TAG_DTYPE[tag]
    rtol = _rtol(dt)
    xr, yr = g[f"{tag}/x"], g[f"{tag}/y"]
    labels = torch.from_numpy(g[f"{tag}/labels"]).to(DEV)
    n = len(labels)
    import item_alignment_b200 as ia
    for m in MEASURES:
        x = torch.from_numpy(xr).to(DEV).to(dt).requires_grad_(True)
        y = torch.from_numpy(yr).to(DEV).to(dt).requires_grad_(True)
        key = f"{tag}/{m}/hinge/0.3"
        sim0, probs0 = F_.pair_score(m, x.detach(), y.detach())
        parity.assert_scores_close(m, sim0, g[key + "/sim"], xr, yr, rtol)
        parity.assert_probs_close(probs0, g[key + "/probs"], rtol, parity.score_atol(m, xr, yr, rtol))
        # reference sequence: head.similarity -> HingeLoss module -> backward, all on our drop-in modules
        sim = ia.PairSimilarity(m)(x, y) if m != "inner_product" else ia.InnerProduct()(x, y)
        assert torch.equal(sim.detach(), sim0)
        loss = ia.HingeLoss(margin=0.3)(sim.view(-1), (labels * 2 - 1).view(-1))
        loss.backward()
        parity.assert_loss_close(loss, g[key + "/loss"], 10 * rtol, 10 * float(parity.score_atol(m, xr, yr, rtol).max()))
        tx, ty = parity.grad_term_scale(m, xr, yr)
        gr = 10 * rtol if dt == torch.float32 else 4e-3      # grads come back in the input dtype (bf16 ulp 2^-8)
        parity.assert_grad_close(x.grad, g[key + "/dx"], tx, 1.0 / n, gr, key + "/dx")
        parity.assert_grad_close(y.grad, g[key + "/dy"], ty, 1.0 / n, gr, key + "/dy")


@pytest.mark.parametrize("dt,n,d", [(torch.bfloat16, 4099, 1024), (torch.float32, 2051, 1024), (torch.float32, 3000, 768),
                                    (torch.float16, 515, 512), (torch.float32, 300, 1030), (torch.bfloat16, 77, 2048),
                                    (torch.float32, 64, 4100), (torch.bfloat16, 1000, 144)])
def test_fused_vs_oracle_on_seeded_inputs(F_, dt, n, d):
    from oracle import torch_port
    gen = torch.Generator().manual_seed(100 + n + d)
    x = torch.tanh(torch.randn(n, d, generator=gen)).to(dt)
    labels = (torch.rand(n, generator=gen) < 0.5).long()
    y = torch.where(labels[:, None] == 1, torch.tanh(x.float() + 0.25 * torch.randn(n, d, generator=gen)),
                    torch.tanh(torch.randn(n, d, generator=gen))).to(dt)
    rtol = _rtol(dt)
    xd, yd, ld = x.to(DEV), y.to(DEV), labels.to(DEV)
    for m, l in (("inner_product", "bce"), ("cosine", "hinge"), ("l1", "hinge"), ("l2", "euclidean"), ("l1", "euclidean"),
                 ("l2", "hinge"), ("cosine", "cosine"), ("inner_product", "hinge")):
        rs, rp, rl, rdx, rdy = torch_port.pair_score_loss_fwd_bwd(m, l, x, y, labels, 1.0)
        sim, probs, loss, dx, dy = F_.pair_score_loss_raw(m, l, xd, yd, ld, 1.0, "mean", grad_dtype=torch.float32)
        parity.assert_scores_close(m, sim, rs, x.float(), y.float(), rtol, f"{m}/{l} sim")
        parity.assert_probs_close(probs, rp, rtol, parity.score_atol(m, x.float(), y.float(), rtol))
        parity.assert_loss_close(loss, rl, 10 * rtol, 10 * float(parity.score_atol(m, x.float(), y.float(), rtol).max()))
        tx, ty = parity.grad_term_scale(m if l != "cosine" else "cosine", x.float(), y.float())
        parity.assert_grad_close(dx, rdx, tx, 1.0 / n, 10 * rtol, f"{m}/{l} dx")
        parity.assert_grad_close(dy, rdy, ty, 1.0 / n, 10 * rtol, f"{m}/{l} dy")


def test_threshold_labels_bit_exact(F_):
    """labels = probs >= thr exactly as numpy evaluates it for the np.arange sweep of finetune_text.py:576-580."""
    from oracle import torch_port
    gen = torch.Generator().manual_seed(3)
    n, d = 50000, 768                        # BASELINE config 1 shape
    x = torch.tanh(torch.randn(n, d, generator=gen))
    y = torch.tanh(x + 0.8 * torch.randn(n, d, generator=gen))
    x[:5] = 0                                # zero rows -> cosine 0 -> prob exactly 0.5
    xd, yd = x.to(DEV), y.to(DEV)
    for m in MEASURES:
        for thr in np.arange(0.1, 1.0, 0.1):
            sim, probs, labels = F_.pair_score(m, xd, yd, threshold=thr)
            # the label rule applied to OUR fp32 probs must be bit-exact ...
            assert np.array_equal(labels.cpu().numpy(), torch_port.threshold_labels(probs, thr))
        # ... and our probs vs the oracle's differ only by fp32 rounding, so labels can only differ where the
        # oracle's prob is within rounding distance of the threshold
        rs = torch_port.similarity(m, x, y)
        rp = torch_port.probs_of(m, rs)
        sim, probs, labels = F_.pair_score(m, xd, yd, threshold=0.5)
        ref_labels = torch_port.threshold_labels(rp, 0.5)
        diff = labels.cpu().numpy() != ref_labels
        near = np.abs(rp.numpy().astype(np.float64) - 0.5) <= 1e-5 * 0.5 + parity.score_atol(m, x, y, 1e-5)
        assert not (diff & ~near).any(), f"{m}: {int((diff & ~near).sum())} labels differ away from the threshold"
        if m == "cosine":
            assert labels[:5].all() and (probs[:5] == 0.5).all()


def test_reductions_empty_single_and_determinism(F_):
    from oracle import torch_port
    gen = torch.Generator().manual_seed(9)
    x = torch.tanh(torch.randn(777, 256, generator=gen))
    y = torch.tanh(torch.randn(777, 256, generator=gen))
    labels = (torch.rand(777, generator=gen) < 0.5).long()
    xd, yd, ld = x.to(DEV), y.to(DEV), labels.to(DEV)
    sim, _, lsum, dxs, _ = F_.pair_score_loss_raw("l2", "hinge", xd, yd, ld, 1.0, "sum")
    _, _, lmean, dxm, _ = F_.pair_score_loss_raw("l2", "hinge", xd, yd, ld, 1.0, "mean")
    _, _, lnone, dxn, _ = F_.pair_score_loss_raw("l2", "hinge", xd, yd, ld, 1.0, "none")
    rs = torch_port.similarity("l2", x, y)
    ref_none = torch_port.hinge_loss(rs, labels * 2 - 1, 1.0, "none")
    np.testing.assert_allclose(lnone.cpu().numpy(), ref_none.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(lsum), float(ref_none.sum()), rtol=1e-5)
    np.testing.assert_allclose(float(lmean), float(ref_none.mean()), rtol=1e-5)
    assert torch.equal(dxs, dxn)
    torch.testing.assert_close(dxm * 777, dxs, rtol=1e-5, atol=1e-7)
    # bit-reproducible scalar loss run to run (fixed-order two-stage reduction, no float atomics)
    runs = {float(F_.pair_score_loss_raw("cosine", "bce", xd, yd, ld)[2]) for _ in range(5)}
    assert len(runs) == 1
    # N = 1
    s1, p1, l1, dx1, _ = F_.pair_score_loss_raw("cosine", "bce", xd[:1], yd[:1], ld[:1])
    r1 = torch_port.pair_score_loss_fwd_bwd("cosine", "bce", x[:1], y[:1], labels[:1])
    np.testing.assert_allclose(float(l1), float(r1[2]), rtol=1e-5)
    # empty batch: torch's mean over nothing is nan, sum is 0
    e = torch.empty((0, 256), device=DEV)
    el = torch.empty((0,), dtype=torch.long, device=DEV)
    assert torch.isnan(F_.pair_score_loss_raw("cosine", "bce", e, e, el)[2])
    assert float(F_.pair_score_loss_raw("cosine", "bce", e, e, el, reduction="sum")[2]) == 0.0
    assert F_.pair_score("l1", e, e)[0].numel() == 0
    with pytest.raises(ValueError, match="Unsupported similarty measure"):
        F_.pair_score("softmax", xd, yd)


def test_non_contiguous_rows_and_upstream_scale(F_):
    """Strided row views (ld > d) go through the same kernels; a non-unit upstream gradient (GradScaler)
    rescales the stored gradients on the device."""
    from oracle import torch_port
    gen = torch.Generator().manual_seed(21)
    big = torch.tanh(torch.randn(200, 3 * 128, generator=gen))
    labels = (torch.rand(200, generator=gen) < 0.5).long()
    bd = big.to(DEV)
    xv, yv = bd[:, :128], bd[:, 256:384]                 # row stride 384 elements
    x, y = big[:, :128].contiguous(), big[:, 256:384].contiguous()
    xg, yg = xv.detach().clone().requires_grad_(True), None
    bd2 = bd.clone().requires_grad_(True)
    sim, probs, loss = F_.pair_score_loss("cosine", "bce", bd2[:, :128], bd2[:, 256:384], labels.to(DEV))
    (loss * 1024.0).backward()
    rs, rp, rl, rdx, rdy = torch_port.pair_score_loss_fwd_bwd("cosine", "bce", x, y, labels)
    parity.assert_scores_close("cosine", sim, rs, x, y, 1e-5)
    torch.testing.assert_close(bd2.grad[:, :128].cpu(), rdx * 1024.0, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(bd2.grad[:, 256:384].cpu(), rdy * 1024.0, rtol=1e-4, atol=1e-6)
    assert float(bd2.grad[:, 128:256].abs().max()) == 0.0


def test_full_size_properties_config2_and_3(F_):
    """BASELINE configs 2 / 3 at full size (65536 x 1024): size-independent properties instead of the oracle."""
    gen = torch.Generator(device=DEV).manual_seed(20221009)
    n, d = 65536, 1024
    for dt, m, l in ((torch.bfloat16, "inner_product", "bce"), (torch.float32, "l1", "hinge"), (torch.float32, "l2", "euclidean")):
        x = torch.tanh(torch.randn(n, d, device=DEV, generator=gen)).to(dt)
        y = torch.tanh(torch.randn(n, d, device=DEV, generator=gen)).to(dt)
        labels = (torch.rand(n, device=DEV, generator=gen) < 0.5).long()
        sim, probs, loss, dx, dy = F_.pair_score_loss_raw(m, l, x, y, labels, 1.0, "mean", grad_dtype=torch.float32)
        # (1) forward-only kernel and fused kernel agree bit for bit
        sim_f, probs_f, _ = F_.pair_score_raw(m, x, y)
        assert torch.equal(sim, sim_f) and torch.equal(probs, probs_f)
        # (2) symmetry: inner(x,y) == inner(y,x) bitwise; distances are symmetric only up to the +eps convention
        if m == "inner_product":
            assert torch.equal(F_.pair_score_raw(m, y, x)[0], sim)
            # (3) linearity in y under exact power-of-two scaling
            assert torch.equal(F_.pair_score_raw(m, x, (y.float() * 2).to(dt))[0], sim * 2)
            # (4) gradient identity dx = (sigmoid(s) - l)/N * y, from the kernel's own sim
            g = (torch.sigmoid(sim) - labels.float()) / n
            torch.testing.assert_close(dx, g[:, None] * y.float(), rtol=2e-5, atol=1e-9)
            torch.testing.assert_close(dy, g[:, None] * x.float(), rtol=2e-5, atol=1e-9)
        else:
            assert torch.equal(dx, -dy)                               # d/dy = -d/dx for l1 / l2
            if m == "l1":
                g = torch.where((1.0 - sim * (labels * 2 - 1).float()) > 0, -(labels * 2 - 1).float(), torch.zeros_like(sim)) / n
                assert torch.equal(dx, g[:, None] * torch.sign(x - y + 1e-6))
            else:
                # |dx_i| = |g_i| for the unit direction (x-y+eps)/s
                t = (labels * 2 - 1).float()
                gmag = torch.where(t > 0, torch.ones_like(sim), 1.0 / (sim * sim)) / n
                torch.testing.assert_close(dx.norm(dim=1), gmag, rtol=1e-4, atol=0)
        # (5) loss equals the mean of the per-pair losses of the 'none' reduction, to fp32 rounding
        lnone = F_.pair_score_loss_raw(m, l, x, y, labels, 1.0, "none", want_grads=False)[2]
        np.testing.assert_allclose(float(loss), float(lnone.double().mean()), rtol=2e-6)
        # (6) idempotence / determinism
        again = F_.pair_score_loss_raw(m, l, x, y, labels, 1.0, "mean", grad_dtype=torch.float32)
        assert torch.equal(again[2], loss) and torch.equal(again[3], dx)


def test_softmax_head_golden_and_reference_weights(golden, F_):
    """TwoTowerClassificationHead + CE against the reference-module fixtures, and the commented weight sets of
    submit/similarity.py:5-18 against the outputs of its numpy body (:19-24)."""
    import item_alignment_b200 as ia
    g = golden("head_golden")
    for tag in ("s", "m"):
        k = f"twotower/{tag}/"
        t = {n: torch.from_numpy(g[k + n]).to(DEV) for n in ("f1", "f2", "w", "b", "labels")}
        logits, probs, loss, dx, dy, dw, db = F_.softmax_head_raw(t["f1"], t["f2"], t["w"], t["b"], t["labels"])
        for ours, name, tol in ((logits, "logits", 1e-5), (probs, "probs", 1e-5), (loss, "loss", 1e-5), (dx, "dx", 1e-4),
                                (dy, "dy", 1e-4), (dw, "dw", 1e-4), (db, "db", 1e-4)):
            ref = g[k + name]
            np.testing.assert_allclose(ours.cpu().numpy(), ref, rtol=tol, atol=tol * max(1e-3, float(np.abs(ref).max())), err_msg=name)
        # drop-in module, unfused ladder (CrossEntropyLoss on logits) and fused forward_with_loss
        h = g[k + "f1"].shape[1]
        head = ia.TwoTowerClassificationHead(h).to(DEV)
        with torch.no_grad():
            head.out_proj.weight.copy_(t["w"]); head.out_proj.bias.copy_(t["b"])
        f1, f2 = t["f1"].clone().requires_grad_(True), t["f2"].clone().requires_grad_(True)
        x, y, lg, pr = head(f1, f2)
        l2 = torch.nn.CrossEntropyLoss()(lg.view(-1, 2), t["labels"].view(-1))
        l2.backward()
        np.testing.assert_allclose(float(l2), float(g[k + "loss"]), rtol=1e-5)
        np.testing.assert_allclose(f1.grad.cpu().numpy(), g[k + "dx"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(head.out_proj.weight.grad.cpu().numpy(), g[k + "dw"], rtol=1e-4, atol=1e-7)
        head.zero_grad()
        f1b, f2b = t["f1"].clone().requires_grad_(True), t["f2"].clone().requires_grad_(True)
        _, _, _, _, l3 = head.forward_with_loss(f1b, f2b, t["labels"])
        l3.backward()
        np.testing.assert_allclose(float(l3), float(g[k + "loss"]), rtol=1e-5)
        np.testing.assert_allclose(f2b.grad.cpu().numpy(), g[k + "dy"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(head.out_proj.bias.grad.cpu().numpy(), g[k + "db"], rtol=1e-4, atol=1e-7)
    s = golden("submit_golden")
    for i in range(4):
        w, b = s[f"softmax/{i}/w"], s[f"softmax/{i}/b"]
        e1, e2 = s[f"softmax/{i}/e1"], s[f"softmax/{i}/e2"]
        _, probs = F_.softmax_head(torch.from_numpy(e1).float().to(DEV), torch.from_numpy(e2).float().to(DEV),
                                   torch.from_numpy(w).float().to(DEV), torch.from_numpy(b).float().to(DEV))
        np.testing.assert_allclose(probs[:, 1].cpu().numpy(), s[f"softmax/{i}/p1"], rtol=2e-5)


@pytest.mark.parametrize("n,h,dt", [
    (5001, 1024, torch.bfloat16),     # two rows per W read, every lane full, odd n: the last group has a clamped row
    (4100, 768, torch.bfloat16),      # three vectors per lane (h = 768)
    (4097, 1000, torch.bfloat16),     # partial last vector
    (4500, 512, torch.float32),       # fp32, two rows per W read
    (4099, 1024, torch.float32),      # fp32, eight vectors per lane, one row per W read
    (300, 1024, torch.float16),       # small batch: one row per iteration
    (4096, 264, torch.float16),
    (8195, 768, torch.bfloat16),      # four rows per W read (small 16-bit rows, n >= 8192), tail group with three clamped rows
    (9001, 264, torch.float16),
])
def test_softmax_head_ce_vs_oracle_all_kernel_variants(n, h, dt, F_):
    """Seeded shapes that reach every softmax-head kernel variant (ring depth, rows per W read, full / partial lanes),
    forward-only and training, against the oracle restatement of TwoTowerClassificationHead + CrossEntropyLoss."""
    from oracle import torch_port
    gen = torch.Generator().manual_seed(n + h)
    x = torch.tanh(torch.randn(n, h, generator=gen)).to(dt)
    y = torch.tanh(torch.randn(n, h, generator=gen)).to(dt)
    w = torch.randn(2, 2 * h, generator=gen) * 0.05
    b = torch.randn(2, generator=gen) * 0.05
    labels = (torch.rand(n, generator=gen) < 0.5).long()
    rl, rp, rloss, rdx, rdy, rdw, rdb = torch_port.softmax_head_ce_fwd_bwd(x.float(), y.float(), w, b, labels)
    logits, probs, loss, dx, dy, dw, db = F_.softmax_head_raw(x.to(DEV), y.to(DEV), w.to(DEV), b.to(DEV), labels.to(DEV))
    term = float((x.float().abs() @ w[:, :h].abs().t() + y.float().abs() @ w[:, h:].abs().t()).max())
    assert float((logits.cpu() - rl).abs().max()) <= 1e-5 * term
    assert float((probs.cpu() - rp).abs().max()) <= 2e-5 * max(1.0, term)
    np.testing.assert_allclose(float(loss), float(rloss), rtol=2e-5)
    gtol = 2e-5 if dt == torch.float32 else (2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -10)
    for ours, ref, name in ((dx, rdx, "dx"), (dy, rdy, "dy")):
        err = (ours.float().cpu() - ref).abs()
        floor = 2.0 ** -24 if dt == torch.float16 else 0.0        # fp16 gradients of a mean over n rows sit in the subnormals
        assert float((err - gtol * ref.abs()).max()) <= 2e-5 * float(ref.abs().max()) + floor, name
    for ours, ref, name in ((dw, rdw, "dw"), (db, rdb, "db")):
        assert float((ours.cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-7, name
    # forward-only kernel: same logits / probabilities
    l2, p2, _, _, _, _, _ = F_.softmax_head_raw(x.to(DEV), y.to(DEV), w.to(DEV), b.to(DEV))
    assert float((l2.cpu() - rl).abs().max()) <= 1e-5 * term
    assert float((p2.cpu() - rp).abs().max()) <= 2e-5 * max(1.0, term)


def test_vecsim_head_golden_and_fused_step(golden):
    import types
    import item_alignment_b200 as ia
    from oracle import torch_port
    g = golden("head_golden")
    for m in MEASURES:
        k = f"vecsim/{m}/"
        h = g[k + "w"].shape[0]
        cfg = types.SimpleNamespace(cls_layers="12", cls_pool="cls", hidden_size=h, classifier_dropout=0.0,
                                    hidden_dropout_prob=0.0, similarity_measure=m, loss_type="hinge", loss_margin=0.5)
        head = ia.VecSimClassificationHead(cfg).to(DEV).eval()
        head.load_state_dict({"dense.weight": torch.from_numpy(g[k + "w"]), "dense.bias": torch.from_numpy(g[k + "b"])})
        f1, f2 = torch.from_numpy(g[k + "f1"]).to(DEV), torch.from_numpy(g[k + "f2"]).to(DEV)
        with torch.no_grad():
            x, y, sim, probs = head(f1, f2)
        np.testing.assert_allclose(x.cpu().numpy(), g[k + "x"], rtol=1e-5, atol=1e-6)
        parity.assert_scores_close(m, sim, g[k + "sim"], g[k + "x"], g[k + "y"], 2e-5)
        parity.assert_probs_close(probs, g[k + "probs"], 2e-5, parity.score_atol(m, g[k + "x"], g[k + "y"], 2e-5))
        # fused step through the dense+tanh projection: gradients reach dense.weight like the reference's autograd
        labels = (torch.arange(f1.shape[0]) % 2).to(DEV)
        out = ia.two_tower_step(head, cfg, f1, f2, labels)
        out["loss"].backward()
        ref_head_w = torch.from_numpy(g[k + "w"]).clone().requires_grad_(True)
        ref_b = torch.from_numpy(g[k + "b"]).clone().requires_grad_(True)
        rx = torch.tanh(torch.nn.functional.linear(torch.from_numpy(g[k + "f1"]), ref_head_w, ref_b))
        ry = torch.tanh(torch.nn.functional.linear(torch.from_numpy(g[k + "f2"]), ref_head_w, ref_b))
        rl = torch_port.loss_ladder("hinge", torch_port.similarity(m, rx, ry), rx, ry, labels.cpu(), 0.5)
        rl.backward()
        np.testing.assert_allclose(float(out["loss"]), float(rl), rtol=1e-5, atol=1e-6)
        gw = head.dense.weight.grad.cpu().numpy()
        np.testing.assert_allclose(gw, ref_head_w.grad.numpy(), rtol=1e-3, atol=1e-5 * max(1.0, float(np.abs(ref_head_w.grad.numpy()).max())))


def test_host_buffer_entry_points(F_):
    """ia_pair_score_host / ia_pair_score_loss_host (HOST pointers, chunked H2D pipeline) agree with the device path."""
    from item_alignment_b200 import _lib
    from item_alignment_b200._lib import check, lib
    gen = torch.Generator().manual_seed(33)
    n, d = 40000, 768
    x = torch.tanh(torch.randn(n, d, generator=gen)).pin_memory()
    y = torch.tanh(torch.randn(n, d, generator=gen)).pin_memory()
    labels = (torch.rand(n, generator=gen) < 0.5).long().pin_memory()
    sim = torch.empty(n).pin_memory(); probs = torch.empty(n).pin_memory(); lab = torch.empty(n, dtype=torch.uint8).pin_memory()
    check(lib().ia_pair_score_host(1, _lib.IA_F32, x.data_ptr(), y.data_ptr(), n, d, sim.data_ptr(), probs.data_ptr(), 0.5, lab.data_ptr(), 0))
    s2, p2, l2 = F_.pair_score("cosine", x.to(DEV), y.to(DEV), threshold=0.5)
    assert torch.equal(sim, s2.cpu()) and torch.equal(probs, p2.cpu()) and torch.equal(lab.bool(), l2.cpu())
    loss = torch.empty(1).pin_memory(); dx = torch.empty_like(x).pin_memory(); dy = torch.empty_like(y).pin_memory()
    check(lib().ia_pair_score_loss_host(1, 0, 1.0, 1, _lib.IA_F32, x.data_ptr(), y.data_ptr(), labels.data_ptr(), n, d,
                                        loss.data_ptr(), dx.data_ptr(), dy.data_ptr(), 0))
    _, _, l3, dx3, dy3 = F_.pair_score_loss_raw("cosine", "bce", x.to(DEV), y.to(DEV), labels.to(DEV))
    np.testing.assert_allclose(float(loss), float(l3), rtol=2e-6)
    assert torch.equal(dx, dx3.cpu()) and torch.equal(dy, dy3.cpu())
    # plugin-style batched compute() on plain numpy lists
    import item_alignment_b200 as ia
    ia.configure("cosine")
    try:
        s = ia.compute_many(x[:100].tolist(), y[:100].tolist())
        np.testing.assert_array_equal(s, sim[:100].numpy())
        assert ia.compute(x[0].tolist(), y[0].tolist()) == float(sim[0])
    finally:
        ia.configure("passthrough")


# ----------------------------------------------------------------------------------------------- round 2
@pytest.mark.parametrize("dt,m,l", [(torch.bfloat16, "inner_product", "bce"),            # BASELINE config 2: ROWS=2 instantiation
                                    (torch.float32, "l1", "hinge"), (torch.float32, "l1", "euclidean"),   # config 3: 4 KB rows
                                    (torch.float32, "l2", "hinge"), (torch.float32, "l2", "euclidean"),
                                    (torch.float32, "cosine", "bce"), (torch.bfloat16, "cosine", "cosine"),
                                    (torch.float32, "cosine", "hinge")])
def test_benchmarked_instantiations_vs_oracle_at_full_size(F_, dt, m, l):
    """The kernels bench.py times -- n >= 16384 selects the adjacent-row (ROWS=2) form for 1.5-3 KB rows, D=1024 fp32 the
    4 KB-row form -- compared DIRECTLY with the oracle (reference head + ladder + backward on the CPU) at the full
    65 536 x 1024 of BASELINE configs 2 and 3, on bench.py's own input recipe.  Gradients in fp32 (tolerance) and in the
    input dtype (must equal the fp32 gradients rounded once)."""
    from oracle import torch_port
    import bench
    n, d = bench.N_PAIRS, bench.DIM
    x, y, labels = bench.make_pairs(torch, torch.device(DEV), bench.SEED + 2000, dt)
    xc, yc, lc = x.cpu(), y.cpu(), labels.cpu()
    rs, rp, rl, rdx, rdy = torch_port.pair_score_loss_fwd_bwd(m, l, xc, yc, lc, 1.0)
    sim, probs, loss, dx, dy = F_.pair_score_loss_raw(m, l, x, y, labels, 1.0, "mean", grad_dtype=torch.float32)
    rtol = _rtol(dt)
    xf, yf = xc.float(), yc.float()
    parity.assert_scores_close(m, sim, rs, xf, yf, rtol, f"{m}/{l} sim")
    parity.assert_probs_close(probs, rp, rtol, parity.score_atol(m, xf, yf, rtol))
    parity.assert_loss_close(loss, rl, 10 * rtol, 10 * float(parity.score_atol(m, xf, yf, rtol).max()))
    tx, ty = parity.grad_term_scale(m if l != "cosine" else "cosine", xf, yf)
    parity.assert_grad_close(dx, rdx, tx, 1.0 / n, 10 * rtol, f"{m}/{l} dx")
    parity.assert_grad_close(dy, rdy, ty, 1.0 / n, 10 * rtol, f"{m}/{l} dy")
    if dt != torch.float32:       # the timed form: gradients in the input dtype
        _, _, loss2, dxl, dyl = F_.pair_score_loss_raw(m, l, x, y, labels, 1.0, "mean")
        assert torch.equal(loss2, loss) and torch.equal(dxl, dx.to(dt)) and torch.equal(dyl, dy.to(dt))
    # the host-buffer entry point bench.py's e2e leg calls (chunked, three streams) returns the same gradients
    if (dt, m, l) == (torch.bfloat16, "inner_product", "bce"):
        from item_alignment_b200 import _lib
        lib = _lib.lib()
        xh, yh, lh = xc.pin_memory(), yc.pin_memory(), lc.pin_memory()
        dxh, dyh = torch.empty_like(xh).pin_memory(), torch.empty_like(yh).pin_memory()
        loss_h = torch.zeros(1)
        _lib.check(lib.ia_pair_score_loss_host(0, 0, 1.0, 1, _lib.IA_BF16, xh.data_ptr(), yh.data_ptr(), lh.data_ptr(), n, d,
                                               loss_h.data_ptr(), dxh.data_ptr(), dyh.data_ptr(), 0))
        assert torch.equal(dxh, dxl.cpu()) and torch.equal(dyh, dyl.cpu())
        parity.assert_loss_close(loss_h[0], rl, 1e-4, 1e-4)


@pytest.mark.parametrize("dt,h", [(torch.bfloat16, 1024), (torch.bfloat16, 768), (torch.float16, 512), (torch.float32, 1024)])
def test_softmax_head_ce_vs_oracle_at_full_size(F_, dt, h):
    """The softmax head + CE training kernels at the benchmarked size (65 536 pairs) against the oracle."""
    from oracle import torch_port
    n = 65536
    gen = torch.Generator().manual_seed(77 + h)
    x = torch.tanh(torch.randn(n, h, generator=gen)).to(dt)
    y = torch.tanh(torch.randn(n, h, generator=gen)).to(dt)
    labels = (torch.rand(n, generator=gen) < 0.5).long()
    w = torch.randn(2, 2 * h, generator=gen) * 0.05
    b = torch.randn(2, generator=gen) * 0.1
    rlog, rprob, rl, rdx, rdy, rdw, rdb = torch_port.softmax_head_ce_fwd_bwd(x, y, w, b, labels)
    logits, probs, loss, dx, dy, dw, db = F_.softmax_head_raw(x.to(DEV), y.to(DEV), w.to(DEV), b.to(DEV), labels.to(DEV),
                                                              grad_dtype=torch.float32)
    rtol = _rtol(dt)
    scale = float((x.float().norm(dim=1).max() + y.float().norm(dim=1).max()) * w.norm(dim=1).max())
    np.testing.assert_allclose(logits.cpu().numpy(), rlog.numpy(), rtol=rtol, atol=rtol * scale)
    np.testing.assert_allclose(probs.cpu().numpy(), rprob.numpy(), rtol=rtol, atol=rtol * scale)
    np.testing.assert_allclose(float(loss), float(rl), rtol=10 * rtol)
    wd = (w[1] - w[0]).abs().max().item()
    np.testing.assert_allclose(dx.cpu().numpy(), rdx.numpy(), rtol=10 * rtol, atol=10 * rtol * wd / n)
    np.testing.assert_allclose(dy.cpu().numpy(), rdy.numpy(), rtol=10 * rtol, atol=10 * rtol * wd / n)
    # dW sums 65 536 terms of either sign in fp32 (block partials in a fixed order): bound by the sum of magnitudes
    mag = float((x.float().abs().mean(0).max())) * 1.0
    np.testing.assert_allclose(dw.cpu().numpy(), rdw.numpy(), rtol=1e-3, atol=2e-5 * mag)
    np.testing.assert_allclose(db.cpu().numpy(), rdb.numpy(), rtol=1e-3, atol=2e-5)


def test_autograd_upstream_scale_fp16_gradscaler_and_second_backward(F_):
    """(a) under GradScaler (finetune_text.py:479-482 with --fp16) the upstream scalar must be applied before
    the gradients are rounded to fp16 -- a mean over 64k pairs makes every unscaled fp16 gradient subnormal; (b) a second
    backward through the same node (retain_graph) must not scale anything twice."""
    from oracle import torch_port
    n, d = 65536, 256
    gen = torch.Generator().manual_seed(5)
    x = torch.tanh(torch.randn(n, d, generator=gen)).to(torch.float16)
    y = torch.tanh(torch.randn(n, d, generator=gen)).to(torch.float16)
    labels = (torch.rand(n, generator=gen) < 0.5).long()
    rs, rp, rl, rdx, rdy = torch_port.pair_score_loss_fwd_bwd("cosine", "bce", x, y, labels)
    scale = 65536.0
    for m, l in (("cosine", "bce"),):
        xd, yd = x.to(DEV).requires_grad_(True), y.to(DEV).requires_grad_(True)
        sim, probs, loss = F_.pair_score_loss(m, l, xd, yd, labels.to(DEV))
        (loss * scale).backward(retain_graph=True)                # scaler.scale(loss).backward()
        ref = (rdx * scale)
        got = xd.grad.float().cpu()
        assert float(got.abs().max()) > 0
        # one rounding of the exactly scaled fp32 gradient: within half an fp16 ulp (2^-11 relative) + fp32 noise
        err = (got - ref).abs()
        bound = ref.abs() * 2.0 ** -10 + 1e-3 * ref.abs().max(dim=1, keepdim=True).values + 1e-7
        assert bool((err <= bound).all()), f"worst {float((err / bound).max()):.3g}x bound"
        assert float((got != 0).float().mean()) > 0.99            # nothing flushed to zero
        g1 = xd.grad.clone()
        (loss * scale).backward()                                 # second backward through the same node
        assert torch.equal(xd.grad, (g1.float() * 2).to(torch.float16))
    # plain backward: the forward launch's buffers are handed over untouched (bitwise the raw kernel's gradients)
    xd, yd = x.to(DEV).requires_grad_(True), y.to(DEV).requires_grad_(True)
    _, _, loss = F_.pair_score_loss("cosine", "bce", xd, yd, labels.to(DEV))
    loss.backward()
    raw = F_.pair_score_loss_raw("cosine", "bce", x.to(DEV), y.to(DEV), labels.to(DEV))
    assert torch.equal(xd.grad, raw[3]) and torch.equal(yd.grad, raw[4])
    # softmax head + CE: same contract
    h = d
    w = (torch.randn(2, 2 * h, generator=gen) * 0.05).to(DEV).requires_grad_(True)
    b = torch.zeros(2, device=DEV, requires_grad=True)
    xd, yd = x.to(DEV).requires_grad_(True), y.to(DEV).requires_grad_(True)
    _, _, loss = F_.softmax_head_ce(xd, yd, w, b, labels.to(DEV))
    (loss * scale).backward(retain_graph=True)
    _, _, _, odx, ody, odw, odb = torch_port.softmax_head_ce_fwd_bwd(x, y, w.detach().cpu(), b.detach().cpu(), labels)
    got = xd.grad.float().cpu()
    ref = odx * scale
    assert bool(((got - ref).abs() <= ref.abs() * 2.0 ** -10 + 1e-3 * ref.abs().max() + 1e-7).all())
    np.testing.assert_allclose(w.grad.cpu().numpy(), (odw * scale).numpy(), rtol=1e-3, atol=1e-3 * float((odw * scale).abs().max()))
    g1, w1 = xd.grad.clone(), w.grad.clone()
    (loss * scale).backward()
    assert torch.equal(xd.grad, (g1.float() * 2).to(torch.float16))
    torch.testing.assert_close(w.grad, w1 * 2, rtol=1e-6, atol=0)


def test_loss_ladder_rejects_unsupported_combinations():
    """a VecSim head with loss_type 'ce' (or an unknown loss) must fail loudly, not return loss=None."""
    import types
    import item_alignment_b200 as ia
    cfg = types.SimpleNamespace(cls_layers="12", cls_pool="cls", hidden_size=64, classifier_dropout=0.0,
                                hidden_dropout_prob=0.0, similarity_measure="cosine", loss_type="ce", loss_margin=1.0)
    head = ia.VecSimClassificationHead(cfg).to(DEV)
    f = torch.randn(8, 64, device=DEV)
    labels = torch.randint(0, 2, (8,), device=DEV)
    with pytest.raises(ValueError, match="loss_type"):
        ia.two_tower_step(head, cfg, f, f, labels)
    cfg.loss_type = "focal"
    with pytest.raises(ValueError, match="loss_type"):
        ia.two_tower_step(head, cfg, f, f, labels)
    out = ia.two_tower_step(head, cfg, f, f, None)      # no labels: no loss, like the reference
    assert out["loss"] is None and out["probs"].shape == (8,)
    # CE labels outside {0,1} (e.g. ignore_index = -100, which this head does not offer) poison the loss with NaN
    tt = ia.TwoTowerClassificationHead(64).to(DEV)
    bad = labels.clone(); bad[3] = -100
    assert torch.isnan(ia.functional.softmax_head_ce(f, f, tt.out_proj.weight, tt.out_proj.bias, bad)[2])
    assert torch.isfinite(ia.functional.softmax_head_ce(f, f, tt.out_proj.weight, tt.out_proj.bias, labels)[2])
