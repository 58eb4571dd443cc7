"""GPU parity of catalog retrieval (all-pairs + per-query top-k) through the C-ABI: exact-arithmetic golden
fixture with engineered ties (bit-exact indices), seeded random catalogs against the stable-sort oracle with
the gap-aware rule of tests/parity.py, shard merge == single index, and full-size properties."""
import os

import numpy as np
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MEASURES = ("inner_product", "cosine", "l1", "l2")


def _desc(m):
    return m in ("inner_product", "cosine")


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16])
def test_exact_fixture_with_ties(golden, dt):
    """Entries are multiples of 1/2 (exact in bf16/fp16), D=32: fp32 sums are order independent, so inner-product
    top-k must be BIT-EXACT including ties -> lower row.  fp32 runs the CUDA-core kernel, bf16/fp16 the tcgen05 one."""
    import item_alignment_b200 as ia
    g = golden("retrieval_golden")
    k = int(g["k"])
    q = torch.from_numpy(g["q"]).to(DEV).to(dt)
    c = torch.from_numpy(g["c"]).to(DEV).to(dt)
    with ia.CatalogIndex(c) as index:
        for m in MEASURES:
            scores, rows = index.topk(q, k, m)
            assert (rows >= 0).all()
            amb = parity.assert_topk_matches(scores, rows, g[f"{m}/scores"], k, _desc(m), tol=2e-6, exact=(m == "inner_product"))
            # ties between duplicate rows (same score bit for bit) must come out in ascending row order
            s, r = scores.cpu().numpy(), rows.cpu().numpy()
            same = s[:, 1:] == s[:, :-1]
            assert (r[:, 1:][same] > r[:, :-1][same]).all(), f"{m}: tie not broken by lower row"
            # sortedness
            d = np.diff(s, axis=1)
            assert (d <= 0).all() if _desc(m) else (d >= 0).all()


@pytest.mark.parametrize("dt,q_n,c_n,d,k", [
    (torch.bfloat16, 300, 5000, 256, 10), (torch.bfloat16, 130, 70000, 1024, 100), (torch.bfloat16, 64, 3001, 72, 128),
    (torch.float16, 257, 4097, 512, 7), (torch.float32, 100, 3000, 96, 20), (torch.bfloat16, 5, 300, 64, 1),
])
def test_random_catalog_vs_oracle(dt, q_n, c_n, d, k):
    import item_alignment_b200 as ia
    from oracle import torch_port
    gen = torch.Generator().manual_seed(q_n + c_n + d)
    cat = torch.tanh(torch.randn(c_n, d, generator=gen)).to(dt)
    q = cat[torch.randint(0, c_n, (q_n,), generator=gen)].float() + 0.1 * torch.randn(q_n, d, generator=gen)
    q = torch.tanh(q).to(dt)
    cat[c_n // 2: c_n // 2 + 50] = cat[:50]          # exact duplicates -> ties
    measures = MEASURES if (dt == torch.float32 or c_n <= 5000) else ("inner_product", "cosine")
    with ia.CatalogIndex(cat.to(DEV)) as index:
        for m in measures:
            scores, rows = index.topk(q.to(DEV), k, m)
            ref = torch.cat([torch_port.all_pairs_scores(m, q[i:i + 64], cat) for i in range(0, q_n, 64)]).numpy()
            scale = float(np.abs(ref).max())
            tol = (2e-6 if dt == torch.float32 else 4e-6) * max(scale, 1.0)
            amb = parity.assert_topk_matches(scores, rows, ref, k, _desc(m), tol=tol)
            assert amb <= 0.2 * q_n * k + 64, f"{m}: too many gap-ambiguous positions ({amb})"
            s, r = scores.cpu().numpy(), rows.cpu().numpy()
            same = s[:, 1:] == s[:, :-1]
            assert (r[:, 1:][same] > r[:, :-1][same]).all(), f"{m}: tie not broken by lower row"


def test_k_larger_than_catalog_and_argument_errors():
    import item_alignment_b200 as ia
    cat = torch.randn(5, 64, device=DEV).to(torch.bfloat16)
    q = torch.randn(3, 64, device=DEV).to(torch.bfloat16)
    with ia.CatalogIndex(cat) as index:
        scores, rows = index.topk(q, 8, "inner_product")
        assert (rows[:, :5] >= 0).all() and (rows[:, 5:] == -1).all() and torch.isinf(scores[:, 5:]).all()
        assert sorted(rows[0, :5].tolist()) == [0, 1, 2, 3, 4]
        with pytest.raises(ValueError):
            index.topk(q, 129, "cosine")
        with pytest.raises(ValueError, match="Unsupported similarty measure"):
            index.topk(q, 4, "softmax")
    with pytest.raises(RuntimeError, match="closed"):
        index.topk(q, 4, "cosine")


@pytest.mark.parametrize("m,dt", [("cosine", torch.bfloat16), ("inner_product", torch.bfloat16), ("l2", torch.float32)])
def test_shard_merge_equals_single_index(m, dt):
    """Row shards with global row bases + ia_topk_merge reproduce the unsharded keys bit for bit (the 1-GPU
    emulation of the 2/4/8-GPU path; same kernels, same merge, only the all-gather is missing)."""
    import item_alignment_b200 as ia
    gen = torch.Generator().manual_seed(77)
    c_n, d, q_n, k = 9001, 128, 150, 25
    cat = torch.tanh(torch.randn(c_n, d, generator=gen)).to(dt)
    cat[7000:7040] = cat[100:140]                    # ties across shard boundaries
    q = cat[torch.randint(0, c_n, (q_n,), generator=gen)].to(DEV)
    catd = cat.to(DEV)
    with ia.CatalogIndex(catd) as index:
        whole = index.topk_keys(q, k, m)
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            lo, hi = ia.shard_bounds(c_n, world, r)
            with ia.CatalogIndex(catd[lo:hi], row_base=lo) as shard:
                parts.append(shard.topk_keys(q, k, m))
        merged = ia.merge_keys(torch.stack(parts), k)
        assert torch.equal(merged, whole), f"world {world}: merged shard keys differ from the single index"
        # the probe protocol of ShardedCatalogIndex, emulated on one GPU: every shard reports the ceil(k/G)-th best of a
        # prefix of its rows, the minimum over shards seeds every shard's thresholds; the result must not change
        kp = -(-k // world)
        words = []
        for r in range(world):
            lo, hi = ia.shard_bounds(c_n, world, r)
            n_probe = max(kp, (hi - lo) // 4)
            with ia.CatalogIndex(catd[lo:lo + n_probe], row_base=lo) as probe:
                pk = probe.topk_keys(q, kp, m)
            words.append((pk[:, kp - 1] >> 32) & 0xFFFFFFFF)
        bound = torch.stack(words).min(dim=0).values
        assert int(bound.min()) > 0
        parts = []
        for r in range(world):
            lo, hi = ia.shard_bounds(c_n, world, r)
            with ia.CatalogIndex(catd[lo:hi], row_base=lo) as shard:
                parts.append(shard.topk_keys(q, k, m, init_tau=bound))
                seeded = shard.last_stats()["appends"]
                shard.topk_keys(q, k, m)
                assert seeded <= shard.last_stats()["appends"]       # the bound can only remove work
        merged = ia.merge_keys(torch.stack(parts), k)
        assert torch.equal(merged, whole), f"world {world}: seeded shard keys differ from the single index"
    scores, rows = ia.unpack_keys(whole, m)
    assert int(rows.max()) < c_n and int(rows.min()) >= 0


def test_merge_and_unpack_vs_oracle_keys():
    import item_alignment_b200 as ia
    from oracle import formula
    rng = np.random.default_rng(5)
    g, q, k = 5, 37, 100
    sc = rng.standard_normal((g, q, 300)).astype(np.float32)
    sc[:, :, :40] = np.round(sc[:, :, :40])            # many equal scores
    idx = rng.permutation(g * 300).reshape(g, 1, 300).repeat(q, 1)
    for desc in (True, False):
        keys = formula.pack_keys(sc, idx, descending=desc)
        parts = np.sort(keys, axis=2)[:, :, ::-1][:, :, :k].copy()
        ours = ia.merge_keys(torch.from_numpy(parts.view(np.int64)).to(DEV), k)
        ref = formula.merge_topk_keys(parts, k)
        assert np.array_equal(ours.cpu().numpy().view(np.uint64), ref)
        s, r = ia.unpack_keys(ours, "cosine" if desc else "l2")
        rs, ri = formula.unpack_keys(ref, descending=desc)
        assert np.array_equal(s.cpu().numpy(), rs) and np.array_equal(r.cpu().numpy(), ri)


def test_full_size_properties_config4_slab():
    """BASELINE config 4 shape (1M x 1024 bf16 catalog, k=100) on a query slab: self-retrieval, sortedness,
    duplicate ties by lower row, and a CPU-oracle spot check on a few queries."""
    import item_alignment_b200 as ia
    from oracle import torch_port
    gen = torch.Generator(device=DEV).manual_seed(20221009 + 4000)
    c_n, d, q_n, k = 1_000_000, 1024, 1024, 100
    cat = torch.empty((c_n, d), dtype=torch.bfloat16, device=DEV)
    for lo in range(0, c_n, 100_000):
        cat[lo:lo + 100_000] = torch.tanh(torch.randn(100_000, d, device=DEV, generator=gen)).to(torch.bfloat16)
    cat[900_000:900_500] = cat[1000:1500]              # exact duplicates
    qidx = torch.randint(0, c_n, (q_n,), device=DEV, generator=gen)
    qidx[:100] = torch.arange(1000, 1100, device=DEV)  # queries that have an exact duplicate
    q = cat[qidx].clone()
    with ia.CatalogIndex(cat) as index:
        scores, rows = index.topk(q, k, "cosine")
        torch.cuda.synchronize()
    assert (scores[:, :-1] >= scores[:, 1:]).all()
    # every query is a catalog row: best hit is itself (or its lower-numbered duplicate) with cosine 1
    expect = qidx.clone()
    dup = (qidx >= 900_000) & (qidx < 900_500)
    expect[dup] = qidx[dup] - 900_000 + 1000
    assert torch.equal(rows[:, 0], expect)
    assert (scores[:, 0] - 1.0).abs().max() < 1e-5
    assert torch.equal(rows[:100, 1], torch.arange(900_000, 900_100, device=DEV))     # the duplicate comes second
    # spot check against the CPU oracle
    sub = slice(0, 8)
    ref = torch_port.all_pairs_scores("cosine", q[sub].cpu(), cat.cpu()).numpy()
    parity.assert_topk_matches(scores[sub], rows[sub], ref, k, True, tol=4e-6)


def _nccl_worker(rank, world, port, out_dir):
    import os
    import sys
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import item_alignment_b200 as ia
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    gen = torch.Generator().manual_seed(123)
    c_n, d, q_n, k = 20011, 256, 200, 50
    cat = torch.tanh(torch.randn(c_n, d, generator=gen)).to(torch.bfloat16)
    cat[15000:15030] = cat[10:40]
    q = cat[torch.randint(0, c_n, (q_n,), generator=gen)]
    lo, hi = ia.shard_bounds(c_n, world, rank)
    dev = torch.device("cuda", rank)
    sharded = ia.ShardedCatalogIndex(cat[lo:hi].to(dev), c_n)
    keys = sharded.topk_keys(q.to(dev), k, "cosine")                      # 200 queries: slice-wise merge (all-to-all + all-gather)
    keys_odd = sharded.topk_keys(q[:131].to(dev), k, "cosine")             # odd count: padded slices
    sharded.slice_merge = False
    keys_ag = sharded.topk_keys(q.to(dev), k, "cosine")                   # one all-gather, every rank merges everything
    # two query halves with the first half's exchange overlapped with the second half's scan (q >= 2048)
    sharded.slice_merge = True
    big_q = cat[torch.randint(0, c_n, (2300,), generator=gen)].to(dev)
    keys_big = sharded.topk_keys(big_q, k, "cosine")
    sharded.overlap = False
    keys_big_serial = sharded.topk_keys(big_q, k, "cosine")
    with ia.CatalogIndex(cat.to(dev)) as whole:
        ref = whole.topk_keys(q.to(dev), k, "cosine")
        ref_big = whole.topk_keys(big_q, k, "cosine")
    ok = (torch.equal(keys, ref) and torch.equal(keys_ag, ref) and torch.equal(keys_odd, ref[:131]) and torch.equal(keys_big, ref_big)
          and torch.equal(keys_big_serial, ref_big))
    sharded.close()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write(str(bool(ok)))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_sharded_retrieval_nccl(tmp_path):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(2)] == ["True", "True"]


def test_single_gpu_probed_topk_equals_plain_topk():
    """ia_catalog_topk runs its own probe pass for k > 16 (1/32 of the catalog on the register top-k path seeds every item's
    threshold): the keys must be exactly those of the plain cold-start scan, and the probe must remove list insertions.
    Also: ia_catalog_probe_bound (the shards' entry point) returns a bound that at least k rows meet."""
    import item_alignment_b200 as ia
    gen = torch.Generator().manual_seed(31)
    cat = torch.tanh(torch.randn(200_000, 128, generator=gen)).bfloat16()
    cat[7000:7100] = cat[:100]
    q = (cat[torch.randint(0, 200_000, (300,), generator=gen)].float() + 0.05 * torch.randn(300, 128, generator=gen)).bfloat16()
    q[:50] = cat[:50]                                    # exact ties with the duplicated rows
    qd = q.to(DEV)
    with ia.CatalogIndex(cat.to(DEV)) as index:
        for measure in ("cosine", "inner_product"):
            for k in (100, 128, 40, 17):
                probed = index.topk_keys(qd, k, measure)
                a_probed = index.last_stats()["appends"]
                plain = index.topk_keys_unprobed(qd, k, measure)
                a_plain = index.last_stats()["appends"]
                assert torch.equal(probed, plain), (measure, k)
                assert a_probed < a_plain, (measure, k, a_probed, a_plain)
            # the shards' probe: 4 groups x top-13 of 2048 rows each; >= 52 >= k = 50 rows meet the bound
            bound = index.probe_bound(qd, 13, 4, 2048, measure)
            keys = index.topk_keys_unprobed(qd, 50, measure)
            assert int(bound.min()) > 0
            assert bool((((keys[:, 49] >> 32) & 0xFFFFFFFF) >= bound).all())
            assert torch.equal(index.topk_keys(qd, 50, measure, init_tau=bound), keys)
    with pytest.raises(ValueError):
        with ia.CatalogIndex(cat[:5000].to(DEV)) as small:
            small.probe_bound(qd, 13, 4, 2048, "cosine")            # 4 x 2048 rows do not fit into 5000


@pytest.fixture
def pair_env():
    old = os.environ.get("IA_RETR_PAIR")
    yield
    if old is None:
        os.environ.pop("IA_RETR_PAIR", None)
    else:
        os.environ["IA_RETR_PAIR"] = old


@pytest.mark.parametrize("dt,q_n,c_n,d,k,measure", [
    (torch.bfloat16, 256, 4096, 64, 10, "inner_product"),      # two query tiles: one pair, register top-k path
    (torch.bfloat16, 300, 5000, 128, 10, "cosine"),            # three tiles: the second pair's peer CTA runs a phantom tile
    (torch.float16, 1000, 70000, 256, 100, "cosine"),          # probe pass + append / merge path, ragged last catalog tile
    (torch.bfloat16, 129, 100000, 512, 32, "inner_product"),   # a one-row second tile
    (torch.bfloat16, 2100, 200000, 1024, 128, "cosine"),       # 17 tiles, k at the list capacity
])
def test_cta_pair_kernel_equals_single_cta_kernel(pair_env, dt, q_n, c_n, d, k, measure):
    """The cta_group::2 form of the tensor-core kernel (IA_RETR_PAIR=1: a pair of CTAs runs one 256 x 256 MMA, each CTA keeps the
    128 x 256 accumulator of its own query tile) must return the keys of the single-CTA form bit for bit -- also when seeded
    -- and both must agree with the oracle."""
    import item_alignment_b200 as ia
    from oracle import torch_port
    gen = torch.Generator().manual_seed(q_n + c_n + d + k)
    cat = torch.tanh(torch.randn(c_n, d, generator=gen)).to(dt)
    q = torch.tanh(cat[torch.randint(0, c_n, (q_n,), generator=gen)].float() + 0.1 * torch.randn(q_n, d, generator=gen)).to(dt)
    cat[c_n // 2: c_n // 2 + 50] = cat[:50]
    qd = q.to(DEV)
    with ia.CatalogIndex(cat.to(DEV)) as index:
        out = {}
        for pair in ("0", "1"):
            os.environ["IA_RETR_PAIR"] = pair
            out[pair] = index.topk_keys(qd, k, measure)
            out[pair + "plain"] = index.topk_keys_unprobed(qd, k, measure)
            torch.cuda.synchronize()
        assert torch.equal(out["0"], out["1"])
        assert torch.equal(out["0plain"], out["1plain"])
        assert torch.equal(out["0"], out["0plain"])
        os.environ["IA_RETR_PAIR"] = "1"
        scores, rows = index.topk(qd, k, measure)
        if q_n * c_n <= 1000 * 70000:
            ref = torch.cat([torch_port.all_pairs_scores(measure, q[i:i + 64], cat) for i in range(0, q_n, 64)]).numpy()
            tol = 4e-6 * max(float(np.abs(ref).max()), 1.0)
            parity.assert_topk_matches(scores, rows, ref, k, True, tol=tol)
