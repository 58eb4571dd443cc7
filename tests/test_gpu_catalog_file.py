"""GPU parity of the rank-4 "next" row: catalog files feeding retrieval, and torchkge-style candidate ranking
(torchkge/torchkge/inference.py:216-246) through the l1 / l2 all-pairs kernel, against vectors produced by the
reference's vendored torchkge (tests/golden/kg_golden.npz) and the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_catalog_file_upload_and_index(tmp_path):
    import item_alignment_b200 as ia
    gen = torch.Generator().manual_seed(2)
    c, d = 20_011, 256                                  # > one 32 MiB staging chunk when fp32? no: exercises odd row counts
    cat = torch.tanh(torch.randn(c, d, generator=gen)).bfloat16()
    q = (cat[torch.randint(0, c, (130,), generator=gen)].float() + 0.05 * torch.randn(130, d, generator=gen)).bfloat16()
    path = tmp_path / "cat.iacat"
    ia.write_catalog(path, cat, [f"id{i}" for i in range(c)])
    with ia.CatalogFile(path) as f:
        dev = f.to_device()
        assert torch.equal(dev.cpu(), cat)
        part = f.to_device(1000, 7001)
        assert torch.equal(part.cpu(), cat[1000:7001])
        with ia.CatalogIndex(cat.to(DEV)) as ref_index, f.index() as file_index:
            ks, kr = ref_index.topk(q.to(DEV), 20, "cosine")
            fs, fr = file_index.topk(q.to(DEV), 20, "cosine")
            assert torch.equal(ks, fs) and torch.equal(kr, fr)
            assert f.id(int(fr[0, 0])) == f"id{int(kr[0, 0])}"
        # a shard of the file reports GLOBAL rows: two halves merged == the whole
        with f.index(0, 9000) as lo, f.index(9000, c) as hi:
            parts = torch.stack([lo.topk_keys(q.to(DEV), 20, "cosine"), hi.topk_keys(q.to(DEV), 20, "cosine")])
            ms, mr = ia.unpack_keys(ia.merge_keys(parts, 20), "cosine")
            assert torch.equal(ms, ks) and torch.equal(mr, kr)


def test_large_upload_crosses_staging_chunks(tmp_path):
    import item_alignment_b200 as ia
    gen = torch.Generator().manual_seed(3)
    cat = torch.randn(70_000, 320, generator=gen)       # 89.6 MB fp32: three 32 MiB staging chunks
    path = tmp_path / "big.iacat"
    ia.write_catalog(path, cat)
    with ia.CatalogFile(path) as f:
        assert torch.equal(f.to_device().cpu(), cat)
        assert torch.equal(f.to_device(12_345, 69_999).cpu(), cat[12_345:69_999])


@pytest.mark.parametrize("kind", ["L1", "L2"])
@pytest.mark.parametrize("missing", ["tails", "heads"])
def test_rank_entities_vs_reference_torchkge(golden, kind, missing):
    import item_alignment_b200 as ia
    g = golden("kg_golden")
    ent, rel = torch.from_numpy(g[f"{kind}/ent_emb"]).to(DEV), torch.from_numpy(g[f"{kind}/rel_emb"]).to(DEV)
    ents, rels = torch.from_numpy(g[f"{kind}/ents"]).to(DEV), torch.from_numpy(g[f"{kind}/rels"]).to(DEV)
    ref_scores = g[f"{kind}/{missing}/scores"]            # [n, n_ent] from the reference's inference_scoring_function
    ref_idx, ref_top = g[f"{kind}/{missing}/top_idx"], g[f"{kind}/{missing}/top_scores"]
    pred, scores = ia.rank_entities(ent, rel, ents, rels, top_k=10, missing=missing, dissimilarity_type=kind)
    pred, scores = pred.cpu().numpy(), scores.cpu().numpy()
    assert pred.shape == ref_idx.shape and scores.dtype == np.float32
    # scores: fp32 tolerance on sums of `dim` terms (different summation order; heads also (c + r) - t vs c - (t - r))
    tol = 2e-5 * np.abs(ref_top).max()
    assert np.abs(scores - ref_top).max() <= tol
    # the reported score is the reference's score of the reported entity ...
    assert np.abs(np.take_along_axis(ref_scores, pred, axis=1) - scores).max() <= tol
    # ... and the ranking is the reference's wherever the reference separates neighbours by more than the tolerance
    gaps_ok = np.abs(np.diff(ref_top, axis=1)) > 2 * tol
    safe = np.concatenate([gaps_ok, np.ones((len(pred), 1), bool)], axis=1) & np.concatenate([np.ones((len(pred), 1), bool), gaps_ok], axis=1)
    kth_gap = np.sort(-ref_scores, axis=1)[:, 10] - np.sort(-ref_scores, axis=1)[:, 9] > 2 * tol
    safe[:, -1] &= kth_gap
    assert np.array_equal(pred[safe], ref_idx[safe])
    assert safe.mean() > 0.9
    # sortedness and exact ties (entities 3, 5, 400 are identical rows): lower index first
    assert (np.diff(scores, axis=1) <= 0).all()
    for row in pred:
        pos = {int(e): i for i, e in enumerate(row)}
        present = [e for e in (3, 5, 400) if e in pos]
        assert [pos[e] for e in present] == sorted(pos[e] for e in present)


def test_dissimilarity_knobs_match_pairwise_distance_kernel():
    import item_alignment_b200 as ia
    from oracle import torch_port
    gen = torch.Generator().manual_seed(4)
    cat = torch.randn(3000, 96, generator=gen)
    q = torch.randn(70, 96, generator=gen)
    with ia.CatalogIndex(cat.to(DEV)) as index:
        for p, m in ((1, "l1"), (2, "l2")):
            s0, r0 = index.topk(q.to(DEV), 15, m)                                          # nn.PairwiseDistance semantics
            s1, r1 = index.topk_dissimilarity(q.to(DEV), 15, p=p, eps=1e-6, squared=False)
            assert torch.equal(s0, s1) and torch.equal(r0, r1)
        d2, r2 = index.topk_dissimilarity(q.to(DEV), 15, p=2)                              # torchkge: squared, eps = 0
        ref = torch_port.kg_dissimilarity("L2", q[:, None, :], cat[None, :, :])
        want, widx = torch.sort(ref, dim=1, stable=True)
        assert float((d2.cpu() - want[:, :15]).abs().max()) <= 2e-5 * float(want[:, :15].max())
        assert (r2.cpu() == widx[:, :15]).float().mean() > 0.99
        with pytest.raises(ValueError):
            index.topk_dissimilarity(q.to(DEV), 15, p=3)
