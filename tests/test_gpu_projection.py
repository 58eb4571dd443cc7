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
    # training mode with dropout active goes through the library path (torch RNG), still on the GPU
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
