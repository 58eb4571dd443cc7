"""GPU parity of the head projection fused in front of the score:
x = tanh(dense(f1)), y = tanh(dense(f2)) of VecSimClassificationHead.forward (reference src/models/base.py:67-75,
eval mode) as one tcgen05 GEMM with bias + tanh in its epilogue, and the variant that also scores the pair.

Tolerances.  The kernel accumulates the GEMM in fp32, applies tanh in fp32 and rounds ONCE to the output type; the
oracle is the reference module's arithmetic in fp32 on the same 16-bit inputs.  Embeddings therefore agree up to the
output rounding: |ours - oracle| <= 1 ulp of the 16-bit type (half an ulp of rounding plus accumulation-order noise that can
flip a rounding decision).  Scores are checked against the oracle evaluated on the embeddings the kernel itself wrote
(fp32 tolerance 1e-5 of the term scale, as for the pair kernels) and bit for bit against the stand-alone pair kernel."""
import numpy as np
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MEASURES = ("inner_product", "cosine", "l1", "l2")


def ulp16(ref, dt):
    """1 ulp of the 16-bit type at |ref| (bf16: 8 bits of mantissa, fp16: 11; fp16 subnormal floor 2^-24)."""
    mant = 8 if dt == torch.bfloat16 else 11
    a = ref.abs().clamp_min(2.0 ** -14 if dt == torch.float16 else 1e-30)
    return torch.exp2(torch.floor(torch.log2(a)) - (mant - 1))


def make_case(n, k, h, dt, seed, scale=None):
    gen = torch.Generator().manual_seed(seed)
    f1 = torch.randn(n, k, generator=gen).to(dt)
    f2 = (f1.float() + 0.5 * torch.randn(n, k, generator=gen)).to(dt)
    w = (torch.randn(h, k, generator=gen) * (scale or 1.0 / np.sqrt(k))).to(dt)
    b = torch.randn(h, generator=gen) * 0.1
    return f1, f2, w, b


def acc_noise(f, w, b=None):
    """fp32 accumulation-order noise of the pre-activation: a few 2^-24 of sum |f||w| (tanh' <= 1 carries it to the output)."""
    s = f.float().abs() @ w.float().abs().t()
    return 1e-6 * (s + (b.abs() if b is not None else 0.0))


def check_embeddings(ours, ref, dt, what, noise=0.0):
    ours = ours.float().cpu()
    err = (ours - ref).abs()
    tol = ulp16(ref, dt) + noise
    bad = err > tol
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} of {bad.numel()} beyond 1 ulp, worst {float((err / tol).max()):.2f} ulp"
    exact = (ours == ref.to(dt).float()).float().mean()
    assert float(exact) > 0.98, f"{what}: only {float(exact):.4f} of the elements equal the correctly rounded oracle"


@pytest.mark.parametrize("n,k,h,dt", [
    (21, 48, 48, torch.bfloat16),          # the golden head shape: partial row block, partial K block, partial column tile
    (1000, 2048, 1024, torch.bfloat16),    # cls_layers of two layers concatenated (base.py:47-49)
    (333, 768, 768, torch.float16),
    (4099, 1024, 1024, torch.bfloat16),
    (130, 264, 136, torch.bfloat16),
])
def test_project_tanh_vs_oracle(n, k, h, dt):
    import item_alignment_b200.functional as F_
    from oracle import torch_port
    f1, f2, w, b = make_case(n, k, h, dt, seed=n + k)
    rx, ry, _, _ = torch_port.vecsim_head("inner_product", f1, f2, w.float(), b)
    x, y = F_.project_tanh_raw(f1.to(DEV), f2.to(DEV), w.to(DEV), b.to(DEV))
    assert x.dtype == dt and x.shape == (n, h)
    check_embeddings(x, rx, dt, "x", acc_noise(f1, w, b))
    check_embeddings(y, ry, dt, "y", acc_noise(f2, w, b))
    # no bias
    rx0 = torch.tanh(f1.float() @ w.float().t())
    x0, _ = F_.project_tanh_raw(f1.to(DEV), f2.to(DEV), w.to(DEV), None)
    check_embeddings(x0, rx0, dt, "x (no bias)", acc_noise(f1, w))


def test_project_tanh_golden_head():
    """The reference's own VecSimClassificationHead outputs (tests/golden/head_golden.npz, fp32 module) against the kernel on
    the bf16-rounded inputs: agreement at bf16 input-rounding level, and exact oracle parity on the rounded inputs."""
    import item_alignment_b200.functional as F_
    from oracle import torch_port
    g = np.load("tests/golden/head_golden.npz")
    for m in MEASURES:
        f1, f2, w, b, gx, gsim = (torch.from_numpy(g[f"vecsim/{m}/{k}"]) for k in ("f1", "f2", "w", "b", "x", "sim"))
        q = lambda t: t.to(torch.bfloat16)
        x, y, sim, probs = F_.project_score(m, q(f1).to(DEV), q(f2).to(DEV), q(w).to(DEV), b.to(DEV), want_embeds=True)
        assert float((x.float().cpu() - gx).abs().max()) < 3e-2          # inputs rounded to 8 bits: ~K^0.5 * 2^-9 * |w||f|
        rx, ry, rsim, rprobs = torch_port.vecsim_head(m, q(f1), q(f2), q(w).float(), b)
        check_embeddings(x, rx, torch.bfloat16, f"{m} x", acc_noise(q(f1), q(w), b))
        ox, oy = x.float().cpu(), y.float().cpu()
        parity.assert_scores_close(m, sim, torch_port.similarity(m, ox, oy), ox, oy, 1e-5)
        assert float((sim.cpu() - gsim).abs().max()) < 0.05 * max(1.0, float(gsim.abs().max()))


@pytest.mark.parametrize("n,k,h,dt", [(1000, 1024, 1024, torch.bfloat16), (257, 768, 768, torch.float16),
                                      (21, 48, 48, torch.bfloat16), (5000, 512, 256, torch.bfloat16)])
def test_project_score_vs_oracle_and_pair_kernel(n, k, h, dt):
    import item_alignment_b200.functional as F_
    from oracle import torch_port
    f1, f2, w, b = make_case(n, k, h, dt, seed=7 * n + h)
    args = (f1.to(DEV), f2.to(DEV), w.to(DEV), b.to(DEV))
    xr, yr = F_.project_tanh_raw(*args)
    for m in MEASURES:
        x, y, sim, probs, labels = F_.project_score(m, *args, threshold=0.5, want_embeds=True)
        assert torch.equal(x, xr) and torch.equal(y, yr)              # same embeddings as the projection-only kernel
        ox, oy = x.float().cpu(), y.float().cpu()
        rs = torch_port.similarity(m, ox, oy)
        parity.assert_scores_close(m, sim, rs, ox, oy, 1e-5)
        parity.assert_probs_close(probs, torch_port.probs_of(m, rs), 1e-5, parity.score_atol(m, ox, oy, 1e-5))
        assert np.array_equal(labels.cpu().numpy(), torch_port.threshold_labels(probs, 0.5))
        # against the stand-alone pair kernel on the written embeddings: same values up to fp32 summation order
        s2, p2, _ = F_.pair_score_raw(m, x, y)
        parity.assert_scores_close(m, sim, s2, ox, oy, 1e-5)
        # embeddings never written: identical scores
        sim_n, probs_n = F_.project_score(m, *args)
        assert torch.equal(sim_n, sim) and torch.equal(probs_n, probs)


def test_project_tanh_backward_vs_autograd():
    import item_alignment_b200.functional as F_
    n, k, h, dt = 300, 256, 128, torch.bfloat16
    f1, f2, w, b = make_case(n, k, h, dt, seed=5)
    gen = torch.Generator().manual_seed(6)
    gx, gy = torch.randn(n, h, generator=gen), torch.randn(n, h, generator=gen)
    a = [t.to(DEV).requires_grad_(True) for t in (f1, f2, w, b)]
    x, y = F_.project_tanh(*a)
    (x.float() * gx.to(DEV)).sum().backward(retain_graph=True)
    (y.float() * gy.to(DEV)).sum().backward()
    r = [t.float().clone().requires_grad_(True) for t in (f1, f2, w, b)]
    rx = torch.tanh(torch.nn.functional.linear(r[0], r[2], r[3]))
    ry = torch.tanh(torch.nn.functional.linear(r[1], r[2], r[3]))
    ((rx * gx).sum() + (ry * gy).sum()).backward()
    for ours, ref, name in zip(a, r, ("df1", "df2", "dw", "db")):
        scale = float(ref.grad.abs().max())
        err = float((ours.grad.float().cpu() - ref.grad).abs().max())
        assert err <= 3e-2 * scale, f"{name}: {err} vs scale {scale}"     # bf16 operands in the library GEMMs of the backward


def test_head_module_uses_fused_projection():
    import types
    import item_alignment_b200 as ia
    from oracle import torch_port
    cfg = types.SimpleNamespace(cls_layers="11,12", cls_pool="cls", hidden_size=256, classifier_dropout=0.1,
                                hidden_dropout_prob=0.1, similarity_measure="cosine")
    head = ia.VecSimClassificationHead(cfg).to(DEV).bfloat16().eval()
    gen = torch.Generator().manual_seed(9)
    f1 = torch.randn(77, 512, generator=gen).to(torch.bfloat16)
    f2 = torch.randn(77, 512, generator=gen).to(torch.bfloat16)
    launches = ia._lib.lib().ia_launch_count()
    with torch.no_grad():
        x, y, sim, probs = head(f1.to(DEV), f2.to(DEV))
    assert ia._lib.lib().ia_launch_count() - launches == 2          # GEMM+score kernel and the [N] finalize
    rx, ry, _, _ = torch_port.vecsim_head("cosine", f1, f2, head.dense.weight.float().cpu(), head.dense.bias.float().cpu())
    check_embeddings(x, rx, torch.bfloat16, "x", acc_noise(f1, head.dense.weight.cpu(), head.dense.bias.float().cpu()))
    ox, oy = x.float().cpu(), y.float().cpu()
    parity.assert_scores_close("cosine", sim, torch_port.similarity("cosine", ox, oy), ox, oy, 1e-5)
    # training mode with dropout active: the fused kernels generate the masks (test_head_train_step_fused_vs_oracle_graph)
    head.train()
    x2, y2, sim2, _ = head(f1.to(DEV), f2.to(DEV))
    assert x2.requires_grad and sim2.requires_grad
    # fused single-pass loss through the fused projection when dropout is off
    head.dropout.p = 0.0
    labels = (torch.rand(77, generator=gen) < 0.5).long().to(DEV)
    _, _, _, _, loss = head.forward_with_loss(f1.to(DEV), f2.to(DEV), labels, "hinge", 0.5)
    loss.backward()
    assert head.dense.weight.grad is not None and bool(torch.isfinite(head.dense.weight.grad).all())


def test_project_fast_tanh_is_opt_in_and_within_one_ulp():
    """fast_tanh=True swaps in the hardware tanh.approx (2^-11 relative): outputs stay within one bf16 ulp (two fp16 ulps) of the
    oracle, but are no longer the correctly rounded value most of the time -- which is why it is never the default."""
    import item_alignment_b200.functional as F_
    from oracle import torch_port
    for dt, ulps in ((torch.bfloat16, 1.0), (torch.float16, 2.0)):
        f1, f2, w, b = make_case(1000, 512, 384, dt, seed=77)
        rx, ry, _, _ = torch_port.vecsim_head("cosine", f1, f2, w.float(), b)
        args = (f1.to(DEV), f2.to(DEV), w.to(DEV), b.to(DEV))
        xa, _ = F_.project_tanh_raw(*args)
        xf, yf = F_.project_tanh_raw(*args, fast_tanh=True)
        err = (xf.float().cpu() - rx).abs()
        assert bool((err <= ulps * ulp16(rx, dt) + acc_noise(f1, w, b)).all())
        assert not torch.equal(xa, xf)                                  # a different rounding here and there: really another path
        check_embeddings(xa, rx, dt, "accurate path", acc_noise(f1, w, b))
        sim_f, _ = F_.project_score("cosine", *args, fast_tanh=True)
        ox, oy = xf.float().cpu(), yf.float().cpu()
        parity.assert_scores_close("cosine", sim_f, torch_port.similarity("cosine", ox, oy), ox, oy, 1e-5)


def test_project_empty_batch():
    import item_alignment_b200.functional as F_
    f = torch.zeros(0, 64, device=DEV, dtype=torch.bfloat16)
    w = torch.randn(32, 64, device=DEV).bfloat16()
    x, y = F_.project_tanh_raw(f, f, w, None)
    assert x.shape == (0, 32) and y.shape == (0, 32)
    sim, probs = F_.project_score("cosine", f, f, w, None)
    assert sim.shape == (0,) and probs.shape == (0,)


def test_project_rejects_what_it_cannot_do():
    import item_alignment_b200.functional as F_
    f = torch.randn(8, 64, device=DEV)
    w = torch.randn(64, 64, device=DEV)
    with pytest.raises(NotImplementedError):
        F_.project_tanh_raw(f, f, w, None)                            # fp32: no TF32 substitution
    fb, wb = f.bfloat16(), w.bfloat16()
    with pytest.raises(ValueError):
        F_.project_score("softmax", fb, fb, wb, None)
    with pytest.raises(ValueError):
        F_.project_tanh_raw(fb[:, :60].contiguous(), fb[:, :60].contiguous(), wb[:, :60].contiguous(), None)   # K % 8 != 0
    with pytest.raises(RuntimeError):
        F_.project_tanh_raw(fb.cpu(), fb.cpu(), wb.cpu(), None)


# ----------------------------------------------------------------------------------------------- round 2: training side
def _oracle_train_graph(f1, f2, w, b, p, seed, step, want="xy"):
    """VecSimClassificationHead.forward in train() mode (reference base.py:67-75) with the masks of csrc/philox.cuh replayed on
    the host: x = drop(tanh(dense(drop(f)))).  fp32 torch graph on the 16-bit-quantised leaves."""
    from oracle import formula
    n, k = f1.shape
    h = w.shape[0]
    s = 1.0 / (1.0 - p) if p > 0 else 1.0
    leaves = [t.float().clone().requires_grad_(True) for t in (f1, f2, w, b)]
    outs, masks = [], []
    for side, f in enumerate(leaves[:2]):
        m_in = torch.from_numpy(formula.philox_keep_mask(n, k, p, seed, step, side)) if p > 0 else torch.ones(n, k, dtype=torch.bool)
        m_out = torch.from_numpy(formula.philox_keep_mask(n, h, p, seed, step, 2 + side)) if p > 0 else torch.ones(n, h, dtype=torch.bool)
        if p > 0:      # the kernel rounds the dropped features once: rounded value, straight-through gradient
            fd = (f * m_in * s).detach().to(f1.dtype).float() + (f * m_in * s - (f * m_in * s).detach())
        else:
            fd = f
        t = torch.tanh(torch.nn.functional.linear(fd, leaves[2], leaves[3]))
        outs.append(t * m_out * s)
        masks.append((m_in, m_out))
    return leaves, outs, masks


def test_dropout_masks_replay_and_rate():
    """ia_dropout_fwd against the host replay of the Philox masks (exact), and the keep rate."""
    import item_alignment_b200.functional as F_
    from oracle import formula
    gen = torch.Generator().manual_seed(2)
    for dt, rows, cols, p in ((torch.bfloat16, 1003, 264, 0.1), (torch.float16, 64, 1024, 0.5), (torch.bfloat16, 7, 8, 0.25)):
        x = torch.randn(rows, cols, generator=gen).to(dt)
        for stream_id, step in ((0, 1), (1, 7)):
            out = F_.dropout_raw(x.to(DEV), p, 1234567890123, step, stream_id).cpu()
            keep = torch.from_numpy(formula.philox_keep_mask(rows, cols, p, 1234567890123, step, stream_id))
            ref = torch.where(keep, (x.float() * (1.0 / (1.0 - p))).to(dt), torch.zeros((), dtype=dt))
            assert torch.equal(out, ref), (dt, rows, cols, p, stream_id)
    big = F_.dropout_raw(torch.ones(8192, 1024, dtype=torch.bfloat16, device=DEV), 0.1, 99, 3, 0)
    rate = float((big != 0).float().mean())
    assert abs(rate - 0.9) < 3 * (0.09 / 8192 / 1024) ** 0.5 + 2e-5, rate          # binomial 3 sigma + the 2^-16 quantisation of p
    assert abs(float(big.float().mean()) - 1.0) < 2e-3                                # unbiased up to bf16(1/0.9)
    assert not torch.equal(big, F_.dropout_raw(torch.ones(8192, 1024, dtype=torch.bfloat16, device=DEV), 0.1, 99, 4, 0))   # next step: new mask


@pytest.mark.parametrize("n,k,h,dt,p", [(300, 256, 128, torch.bfloat16, 0.1), (300, 264, 136, torch.bfloat16, 0.0),
                                        (300, 264, 136, torch.bfloat16, 0.1), (1000, 1024, 1024, torch.bfloat16, 0.1),
                                        (129, 768, 768, torch.float16, 0.3), (64, 64, 64, torch.bfloat16, 0.0)])
def test_project_train_forward_and_backward_vs_oracle_graph(n, k, h, dt, p):
    """Training-mode projection: forward (input dropout kernel, GEMM with bias + tanh + output dropout epilogue) and backward
    (tanh / dropout backward, data-gradient GEMM with the input mask in its epilogue, MN-major split-K weight-gradient GEMM,
    bias gradient) against the fp32 autograd graph of the reference's ops with the same masks."""
    import item_alignment_b200.functional as F_
    f1, f2, w, b = make_case(n, k, h, dt, seed=n + h)
    seed, step = 424242, 5
    gen = torch.Generator().manual_seed(n)
    gx, gy = torch.randn(n, h, generator=gen).to(dt), torch.randn(n, h, generator=gen).to(dt)
    a = [t.to(DEV).requires_grad_(True) for t in (f1, f2, w, b)]
    x, y = F_.project_tanh_train(a[0], a[1], a[2], a[3], p, seed, step)
    leaves, (rx, ry), masks = _oracle_train_graph(f1, f2, w, b, p, seed, step)
    for ours, ref, (m_in, m_out), name in ((x, rx, masks[0], "x"), (y, ry, masks[1], "y")):
        o = ours.float().cpu()
        assert bool(((o == 0) | m_out).all()), f"{name}: a dropped element is not zero"
        check_embeddings(ours, ref.detach(), dt, name, acc_noise(f1, w, b) * (1.0 / (1.0 - p)) ** 2 * 2)
    torch.autograd.backward([x, y], [gx.to(DEV), gy.to(DEV)])
    torch.autograd.backward([rx, ry], [gx.float(), gy.float()])
    bad = []
    for ours, ref, name in zip(a, leaves, ("df1", "df2", "dw", "db")):
        scale = float(ref.grad.abs().max())
        err = float((ours.grad.float().cpu() - ref.grad).abs().max())
        if not err <= 3e-2 * scale:                                          # 16-bit d_pre / operands, fp32 accumulation
            bad.append(f"{name}: {err:.4g} vs scale {scale:.4g}")
    assert not bad, "; ".join(bad)
    if p > 0:       # the data gradient is exactly zero where the input was dropped
        assert bool(((a[0].grad.cpu() == 0) | masks[0][0]).all()) and bool(((a[1].grad.cpu() == 0) | masks[1][0]).all())


def test_head_train_step_fused_vs_oracle_graph():
    """two_tower_step in train() mode with dropout 0.1: projection + pair loss + the whole backward on this library's kernels,
    against the reference's ops with replayed masks; GradScaler-style upstream scale and second backward included."""
    import types
    import item_alignment_b200 as ia
    from oracle import torch_port
    cfg = types.SimpleNamespace(cls_layers="12", cls_pool="cls", hidden_size=256, classifier_dropout=None, hidden_dropout_prob=0.1,
                                similarity_measure="cosine", loss_type="bce", loss_margin=1.0)
    torch.manual_seed(77)
    head = ia.VecSimClassificationHead(cfg).to(DEV).bfloat16().train()
    gen = torch.Generator().manual_seed(3)
    n = 515
    f1 = torch.randn(n, 256, generator=gen).to(torch.bfloat16)
    f2 = torch.randn(n, 256, generator=gen).to(torch.bfloat16)
    labels = (torch.rand(n, generator=gen) < 0.5).long()
    a1, a2 = f1.to(DEV).requires_grad_(True), f2.to(DEV).requires_grad_(True)
    before = ia.launch_count()
    out = ia.two_tower_step(head, cfg, a1, a2, labels.to(DEV))
    (out["loss"] * 8.0).backward(retain_graph=True)
    launched = ia.launch_count() - before
    step = head._dropout_step
    w, b = head.dense.weight.detach().cpu(), head.dense.bias.detach().float().cpu()
    leaves, (rx, ry), masks = _oracle_train_graph(f1, f2, w, b, 0.1, torch.initial_seed(), step)
    rsim = torch_port.similarity("cosine", rx, ry)
    rloss = torch_port.loss_ladder("bce", rsim, rx, ry, labels)
    (rloss * 8.0).backward()
    assert abs(float(out["loss"]) - float(rloss)) <= 2e-2 * abs(float(rloss))
    for ours, ref, name in ((a1.grad, leaves[0].grad, "df1"), (a2.grad, leaves[1].grad, "df2"), (head.dense.weight.grad, leaves[2].grad, "dw"),
                            (head.dense.bias.grad, leaves[3].grad, "db")):
        scale = float(ref.abs().max())
        err = float((ours.float().cpu() - ref).abs().max())
        assert err <= 5e-2 * scale, f"{name}: {err} vs scale {scale}"
    g1 = a1.grad.clone()
    (out["loss"] * 8.0).backward()                  # second backward through the same node: gradients double, nothing scaled twice
    torch.testing.assert_close(a1.grad.float(), 2 * g1.float(), rtol=2e-2, atol=2e-2 * float(g1.abs().max()))
    assert launched >= 7        # dropout x2, GEMM, pair launch (+ its no-op re-issue), transpose, dgrad, wgrad + reduce + column sums
    # eval mode: no dropout, inference path untouched
    head.eval()
    with torch.no_grad():
        xe, ye, _, _ = head(f1.to(DEV), f2.to(DEV))
    assert float((xe == 0).float().mean()) < 0.01
