"""CPU: the C-ABI library loads and exports every symbol include/ia_b200.h declares; host-side logic
(argument checks, sharding plan, world_size-2 gloo all-gather + merge) works without a GPU."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ia_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ia_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from item_alignment_b200 import _lib
    names = _declared()
    assert len(names) >= 18
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/ia_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signatures and the header disagree"
    lib = _lib.lib()
    assert b"sm_100a" in lib.ia_version()
    assert lib.ia_workspace_bytes() >= 64
    assert lib.ia_softmax_head_workspace_bytes(1024) > lib.ia_workspace_bytes()


def test_enums_match_header():
    from item_alignment_b200 import _lib
    src = open(os.path.join(ROOT, "include", "ia_b200.h")).read()
    for name, val in (("IA_INNER", 0), ("IA_COSINE", 1), ("IA_L1", 2), ("IA_L2", 3), ("IA_LOSS_BCE", 0),
                      ("IA_LOSS_HINGE", 1), ("IA_LOSS_EUCLIDEAN", 2), ("IA_LOSS_COSINE", 3), ("IA_F32", 0),
                      ("IA_BF16", 1), ("IA_F16", 2), ("IA_RED_NONE", 0), ("IA_RED_MEAN", 1), ("IA_RED_SUM", 2)):
        assert re.search(rf"\b{name} = {val}\b", src), name
    assert _lib.MEASURES == {"inner_product": 0, "cosine": 1, "l1": 2, "l2": 3}
    assert _lib.LOSSES == {"bce": 0, "hinge": 1, "euclidean": 2, "cosine": 3}
    assert int(re.search(r"#define IA_MAX_K (\d+)", src).group(1)) == _lib.IA_MAX_K


def test_no_cpu_fallback_and_reference_error_types():
    import types
    import item_alignment_b200 as ia
    x, y = torch.randn(4, 8), torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ia.functional.pair_score("cosine", x, y)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ia.CatalogIndex(torch.randn(10, 8))
    cfg = types.SimpleNamespace(cls_layers="12", cls_pool="cls", hidden_size=8, classifier_dropout=None,
                                hidden_dropout_prob=0.1, similarity_measure="softmax")
    with pytest.raises(ValueError, match="Unsupported similarty measure"):     # reference base.py:63-64 (sic)
        ia.VecSimClassificationHead(cfg)
    cfg.similarity_measure = "cosine"
    head = ia.VecSimClassificationHead(cfg)
    assert sorted(head.state_dict()) == ["dense.bias", "dense.weight"]          # checkpoint keys of the reference
    assert head.dropout.p == 0.1
    cfg.similarity_measure = "inner_product"
    head = ia.VecSimClassificationHead(cfg)
    assert isinstance(head.similarity, ia.InnerProduct) and hasattr(head, "sigmoid")
    tt = ia.TwoTowerClassificationHead(16, dropout=0.0, num_labels=2)
    assert tt.out_proj.weight.shape == (2, 32)                                  # read by finetune_text.py:712
    assert ia.HingeLoss(margin=0.3).margin == 0.3 and ia.EuclideanDistanceLoss().reduction == "mean"
    assert ia.compute([1.0, 2.0], [3.0, 4.0]) == 3.0                            # submit/similarity.py:27-28


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    from item_alignment_b200 import _lib
    lib = _lib.lib()
    assert lib.ia_pair_score_fwd(7, 0, None, None, 4, 8, 8, 8, None, None, 0.5, None, None) == -1
    assert b"Unsupported similarty measure" in lib.ia_last_error()
    assert lib.ia_pair_score_fwd(0, 9, None, None, 4, 8, 8, 8, None, None, 0.5, None, None) == -2
    assert lib.ia_pair_score_fwd(0, 0, None, None, 4, 8, 4, 8, None, None, 0.5, None, None) == -1   # ld < d
    assert lib.ia_pair_score_fwd(0, 0, None, None, 0, 8, 8, 8, None, None, 0.5, None, None) == 0    # empty batch is fine
    assert lib.ia_topk_merge(None, 1, 4, 4, None, None) == -1


def test_shard_bounds_cover_rows_exactly():
    from item_alignment_b200 import shard_bounds
    for total, world in ((1_000_000, 8), (100_000_000, 8), (10, 4), (3, 8), (1001, 2)):
        spans = [shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(hi - lo <= -(-total // world) for lo, hi in spans)


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from item_alignment_b200 import all_gather_keys, shard_bounds
    from oracle import formula
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    q, c, d, k = 9, 203, 16, 7
    cat = rng.integers(-2, 3, size=(c, d)).astype(np.float32) / 2
    cat[50:60] = cat[150:160]                                  # ties across shards
    qs = cat[rng.integers(0, c, size=q)]
    lo, hi = shard_bounds(c, world, rank)
    sc = qs @ cat[lo:hi].T
    idx = np.broadcast_to(np.arange(lo, hi), sc.shape)
    local = np.sort(formula.pack_keys(sc, idx, descending=True), axis=1)[:, ::-1][:, :k].copy()
    gathered = all_gather_keys(torch.from_numpy(local.view(np.int64)))          # product host logic under gloo
    merged = formula.merge_topk_keys(gathered.numpy().view(np.uint64), k)
    s, i = formula.unpack_keys(merged)
    rs, ri = formula.topk_stable(qs @ cat.T, k)
    ok = np.array_equal(i, ri) and np.array_equal(s, rs.astype(np.float32))
    # slice-wise merge: all-to-all hands every rank the lists of its slice of the (padded) queries
    from item_alignment_b200 import all_to_all_keys
    per = -(-q // world)
    padded = np.zeros((per * world, k), dtype=np.uint64)
    padded[:q] = local
    mine = all_to_all_keys(torch.from_numpy(padded.view(np.int64))).numpy().view(np.uint64)      # [world, per, k]
    merged_slice = formula.merge_topk_keys(mine, k)
    ok = ok and np.array_equal(merged_slice[: max(0, min(per, q - rank * per))], merged[rank * per:(rank + 1) * per])
    open(os.path.join(out_dir, f"ok{rank}"), "w").write(str(bool(ok)))
    dist.destroy_process_group()


def test_world_size_2_gloo_shard_gather_merge(tmp_path):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(2)] == ["True", "True"]
