"""Catalog-scale same-item retrieval: all-pairs scores + per-query top-k, single GPU and row-sharded.

No reference implementation exists (the README only motivates it, README.md:12,16); semantics are the
reference's pairwise similarity (src/models/base.py:54-62) for every (query, catalog row), top-k by a stable
sort -- ties go to the lower catalog row (nearest precedent: torchkge/torchkge/inference.py:243-246).

Results travel as packed 64-bit keys (score goodness << 32 | ~row) so that one unsigned compare orders
candidates identically inside the kernel epilogue, in the shard merge and after the NCCL all-gather.
"""
import ctypes

import torch

from . import _lib
from ._lib import MEASURES, check, lib
from .functional import _DT, _ld, _stream


def shard_bounds(total_rows, world_size, rank):
    """Contiguous row shard [lo, hi) of `rank` (shard r holds rows [r*ceil(C/G), ...))."""
    per = -(-total_rows // world_size)
    lo = min(rank * per, total_rows)
    return lo, min(lo + per, total_rows)


class CatalogIndex:
    """Device-resident catalog [C, D] (borrowed) + inverse norms + TMA descriptor + scratch, behind an
    opaque C handle.  `row_base` is the global id of local row 0 (shard offset)."""

    def __init__(self, catalog, row_base=0):
        if not catalog.is_cuda:
            raise RuntimeError("item_alignment_b200 runs on CUDA tensors only (no CPU fallback)")
        if catalog.dim() != 2 or catalog.dtype not in _DT:
            raise ValueError("catalog must be a [C, D] fp32 / bf16 / fp16 CUDA tensor")
        if catalog.stride(1) != 1:
            catalog = catalog.contiguous()
        self.catalog = catalog              # keep the borrowed memory alive
        self.row_base = int(row_base)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(catalog.device):
            check(lib().ia_catalog_create(ctypes.byref(self._h), _DT[catalog.dtype], catalog.data_ptr(), catalog.shape[0],
                                          catalog.shape[1], _ld(catalog), self.row_base, _stream()))

    def close(self):
        if self._h:
            torch.cuda.synchronize(self.catalog.device)
            lib().ia_catalog_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def topk_keys(self, queries, k, measure="cosine", init_tau=None):
        """[Q, k] uint64 keys (as int64 storage), best first.  init_tau: optional int64 [Q] lower bounds on the final
        k-th best key's upper word (see ia_catalog_topk_seeded); rows below them are not collected."""
        if measure not in MEASURES:
            raise ValueError(f"Unsupported similarty measure: {measure}")
        if not self._h:
            raise RuntimeError("CatalogIndex is closed")
        if not queries.is_cuda or queries.dim() != 2 or queries.shape[1] != self.catalog.shape[1]:
            raise ValueError("queries must be a [Q, D] CUDA tensor with the catalog's D")
        if queries.dtype != self.catalog.dtype:
            queries = queries.to(self.catalog.dtype)
        if queries.stride(1) != 1:
            queries = queries.contiguous()
        keys = torch.empty((queries.shape[0], k), dtype=torch.int64, device=queries.device)
        if init_tau is not None:
            init_tau = init_tau.to(torch.int64).contiguous()
            if init_tau.numel() != queries.shape[0]:
                raise ValueError("init_tau needs one entry per query")
        with torch.cuda.device(queries.device):
            check(lib().ia_catalog_topk_seeded(self._h, MEASURES[measure], queries.data_ptr(), queries.shape[0], _ld(queries),
                                               int(k), init_tau.data_ptr() if init_tau is not None else None,
                                               keys.data_ptr(), _stream()))
        return keys

    def topk(self, queries, k, measure="cosine"):
        """(scores [Q,k] fp32, rows [Q,k] int64 global ids), best first; -1 / +-inf where fewer than k rows exist."""
        return unpack_keys(self.topk_keys(queries, k, measure), measure)

    def probe_bound(self, queries, kp, groups, rows_per_group, measure="cosine"):
        """int64 [Q] lower bounds from a probe pass over `groups` disjoint row groups (rows_per_group rows each, a multiple of
        256) at the head of this catalog, top-kp (<= 16) per group: the smallest of the groups' kp-th best key words, 0 = none
        (ia_catalog_probe_bound).  ia_catalog_topk runs this on its own for k > 16; shards call it to agree on a bound."""
        if measure not in ("cosine", "inner_product"):
            raise ValueError("probe passes exist for the tensor-core measures (cosine, inner_product)")
        if queries.dtype != self.catalog.dtype:
            queries = queries.to(self.catalog.dtype)
        if queries.stride(1) != 1:
            queries = queries.contiguous()
        out = torch.empty(queries.shape[0], dtype=torch.int64, device=queries.device)
        with torch.cuda.device(queries.device):
            check(lib().ia_catalog_probe_bound(self._h, MEASURES[measure], queries.data_ptr(), queries.shape[0], _ld(queries), int(kp),
                                               int(groups), int(rows_per_group), out.data_ptr(), _stream()))
        return out

    def topk_keys_unprobed(self, queries, k, measure="cosine"):
        """topk_keys without the library's own probe pass (a zero bound carries no information but counts as seeded): the
        plain cold-start scan, kept for A/B measurements and for the test that both give the same keys."""
        return self.topk_keys(queries, k, measure, init_tau=torch.zeros(queries.shape[0], dtype=torch.int64, device=queries.device))

    def topk_dissimilarity(self, queries, k, p=2, eps=0.0, squared=None):
        """The k SMALLEST l1 / l2 dissimilarities per query with explicit eps / squaring: (dist [Q,k] ascending, rows).
        Defaults are torchkge's plain norms (torchkge/utils/dissimilarities.py:11-25: l1, l2 squared);
        eps=1e-6, squared=False is nn.PairwiseDistance."""
        if p not in (1, 2):
            raise ValueError("p must be 1 or 2")
        squared = (p == 2) if squared is None else bool(squared)
        if not queries.is_cuda or queries.dim() != 2 or queries.shape[1] != self.catalog.shape[1]:
            raise ValueError("queries must be a [Q, D] CUDA tensor with the catalog's D")
        queries = queries.to(self.catalog.dtype)
        if queries.stride(1) != 1:
            queries = queries.contiguous()
        keys = torch.empty((queries.shape[0], k), dtype=torch.int64, device=queries.device)
        with torch.cuda.device(queries.device):
            check(lib().ia_catalog_topk_dissimilarity(self._h, int(p), float(eps), int(squared), queries.data_ptr(),
                                                      queries.shape[0], _ld(queries), int(k), keys.data_ptr(), _stream()))
        return unpack_keys(keys, "l2")

    def last_stats(self):
        """Telemetry of the last topk call (synchronises): appended keys, merges, rare groups, rare blocks."""
        out = (ctypes.c_uint64 * 8)()
        check(lib().ia_catalog_last_stats(self._h, out))
        sp, tp = ctypes.c_int(0), ctypes.c_int(0)
        check(lib().ia_catalog_last_plan(self._h, ctypes.byref(sp), ctypes.byref(tp)))
        return dict(appends=out[0], compactions=out[1], rare_groups=out[2], rare_blocks=out[3], cyc_wait=out[4],
                    cyc_compact=out[5], cyc_total=out[6], cyc_rare=out[7], splits=sp.value, tiles_per_split=tp.value)


def unpack_keys(keys, measure):
    descending = measure in ("inner_product", "cosine")
    scores = torch.empty(keys.shape, dtype=torch.float32, device=keys.device)
    rows = torch.empty(keys.shape, dtype=torch.int64, device=keys.device)
    with torch.cuda.device(keys.device):
        check(lib().ia_unpack_keys(keys.data_ptr(), keys.numel(), int(descending), scores.data_ptr(), rows.data_ptr(), _stream()))
    return scores, rows


def merge_keys(parts, k):
    """[G, Q, k] sorted key lists -> [Q, k] (the merge after the shard all-gather)."""
    parts = parts.contiguous()
    g, q, kk = parts.shape
    if kk != k:
        raise ValueError("merge_keys expects lists of length k")
    out = torch.empty((q, k), dtype=torch.int64, device=parts.device)
    with torch.cuda.device(parts.device):
        check(lib().ia_topk_merge(parts.data_ptr(), g, q, k, out.data_ptr(), _stream()))
    return out


def all_gather_keys(keys, group=None):
    """ONE collective per query batch: all-gather of the per-shard key lists [Q, k] -> [G, Q, k]
    (NCCL over NVLink on the GPU box; gloo in the CPU tests of the host logic)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    keys = keys.contiguous()
    # concatenated (not stacked) output layout: the form both NCCL and gloo accept
    out = torch.empty((world * keys.shape[0],) + tuple(keys.shape[1:]), dtype=keys.dtype, device=keys.device)
    dist.all_gather_into_tensor(out, keys, group=group)
    return out.view((world,) + tuple(keys.shape))


def all_to_all_keys(keys, group=None):
    """First half of the slice-wise merge: rank r receives, from every rank, the key lists of ITS slice of the queries.
    keys [Q, k] with Q a multiple of the world size -> [G, Q/G, k] (entry g = rank g's lists for this rank's queries)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    q, k = keys.shape
    if q % world:
        raise ValueError("all_to_all_keys needs the query count padded to a multiple of the world size")
    send = keys.contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv.view(world, q // world, k)


class ShardedCatalogIndex:
    """Row-sharded retrieval over the ranks of a torch.distributed group (one process per GPU).

    Each rank holds catalog rows [lo, hi) of the global catalog; queries are replicated; every rank computes its
    local top-k as global-row keys, ONE all-gather ([Q,k] u64 per rank) moves them over NVLink, and every rank
    merges the G lists with the same unsigned compare, so ties still break by global row."""

    def __init__(self, local_catalog, total_rows, group=None, probe_fraction=1.0 / 16, slice_merge=None, overlap=False):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.total_rows = int(total_rows)
        lo, hi = shard_bounds(self.total_rows, self.world, self.rank)
        if local_catalog.shape[0] != hi - lo:
            raise ValueError(f"rank {self.rank} must hold rows [{lo}, {hi}) of the catalog, got {local_catalog.shape[0]} rows")
        self.lo, self.hi = lo, hi
        self.local = CatalogIndex(local_catalog, row_base=lo) if hi > lo else None
        # probe = disjoint row groups at the head of the local shard, scanned first with a small k' to agree on a
        # global lower bound (probe_bound)
        self.probe_fraction = probe_fraction
        # slice-wise merge pays from 8 ranks on (measured, C4: 8 GPUs 2.44 vs 2.53 ms; 4 GPUs 4.63 vs 4.60 ms)
        self.slice_merge = (self.world >= 8) if slice_merge is None else bool(slice_merge)
        self.overlap = overlap
        self._side = None

    def _probe_plan(self, k):
        """(kp, groups, rows_per_group) of this rank's probe, or None.  k' = ceil(k / (G*g)) <= 16 keeps the probe on the
        kernel's register top-k path; the groups cover probe_fraction of the local shard."""
        if self.local is None or self.world == 1 or self.probe_fraction <= 0:
            return None
        rows = self.hi - self.lo
        groups = max(1, -(-k // (16 * self.world)))
        kp = -(-k // (self.world * groups))
        per = max(2048, int(rows * self.probe_fraction) // groups) // 256 * 256
        if per * groups > rows or self.local.catalog.dtype == torch.float32:
            return None
        return kp, groups, per

    def close(self):
        if self.local is not None:
            self.local.close()

    def probe_bound(self, queries, k, measure):
        """Lower bound on every query's final k-th best key word, agreed by all ranks with ONE small all-reduce.
        Every probe group (G ranks x g groups, disjoint rows) reports its k'-th best, k' = ceil(k / (G*g)); the
        minimum over all groups is met or exceeded by at least G*g*k' >= k catalog rows, so nothing below it can be
        in the global top-k.  One launch per rank (ia_catalog_probe_bound)."""
        plan = self._probe_plan(k) if measure in ("cosine", "inner_product") else None
        if plan is not None:
            words = self.local.probe_bound(queries, plan[0], plan[1], plan[2], measure)
        else:                      # no probe on this rank (tiny, empty or fp32 shard): no information, no bound
            words = torch.zeros(queries.shape[0], dtype=torch.int64, device=queries.device)
        self.dist.all_reduce(words, op=self.dist.ReduceOp.MIN, group=self.group)
        return words

    def gather_keys(self, keys):
        return all_gather_keys(keys, self.group)

    def _exchange_and_merge(self, keys, k):
        """Per-shard key lists [q, k] -> the merged global lists [q, k] on every rank."""
        q = keys.shape[0]
        if not self.slice_merge or q < 64 * self.world:
            return merge_keys(self.gather_keys(keys), k)             # one all-gather, every rank merges everything
        # Slice-wise merge: an all-to-all hands rank r the G lists of its 1/G of the queries, it merges them, and an
        # all-gather of the merged slices gives every rank the full result.  2 x q*k*8 bytes per rank on the wire instead of
        # G x, and 1/G of the merge work (C4 at 8 GPUs: 16 MB instead of 64 MB per rank).  Same keys: the merge is the same
        # unsigned max-compare, only distributed.
        per = -(-q // self.world)
        if per * self.world != q:
            keys = torch.cat((keys, torch.zeros((per * self.world - q, k), dtype=keys.dtype, device=keys.device)))
        mine = merge_keys(all_to_all_keys(keys, self.group), k)      # [per, k]
        out = torch.empty((per * self.world, k), dtype=keys.dtype, device=keys.device)
        self.dist.all_gather_into_tensor(out, mine, group=self.group)
        return out[:q]

    def topk_keys(self, queries, k, measure="cosine", use_probe=True):
        bound = None
        if self.world > 1 and use_probe:
            bound = self.probe_bound(queries, k, measure)
        q = queries.shape[0]
        if self.world == 1:
            return self.local.topk_keys(queries, k, measure, init_tau=bound)

        def local_scan(a, b):
            if self.local is None:
                return torch.zeros((b - a, k), dtype=torch.int64, device=queries.device)
            return self.local.topk_keys(queries[a:b], k, measure, init_tau=None if bound is None else bound[a:b])

        if not self.overlap or q < 2048:
            return self._exchange_and_merge(local_scan(0, q), k)
        # (opt-in, overlap=True) Two query halves (cut at a multiple of the 128-query tile): the exchange + merge of the first
        # half runs on a side stream while the tensor cores scan the second half.  Measured SLOWER (C4, 2 GPUs: 9.31 vs 8.88 ms):
        # the scan is a persistent kernel with one CTA per SM, and NCCL's kernels take SMs away from it -- some CTAs of the
        # second scan start a wave late.  Kept for workloads whose scan does not fill the GPU.
        cut = (q // 2 + 127) // 128 * 128
        out = torch.empty((q, k), dtype=torch.int64, device=queries.device)
        main = torch.cuda.current_stream(queries.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=queries.device)
        side = self._side
        for a, b in ((0, cut), (cut, q)):
            part = local_scan(a, b)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(side):
                side.wait_event(done)
                out[a:b] = self._exchange_and_merge(part, k)
                part.record_stream(side)
        out.record_stream(side)
        main.wait_stream(side)
        return out

    def topk(self, queries, k, measure="cosine"):
        return unpack_keys(self.topk_keys(queries, k, measure), measure)


def rank_entities(ent_emb, rel_emb, known_entities, known_relations, top_k=1, missing="tails", dissimilarity_type="L2",
                  index=None):
    """TransE candidate ranking of torchkge's EntityInference (torchkge/torchkge/inference.py:216-246) through the
    catalog kernels: scores = -dissimilarity(h + r, c) (tails) or -dissimilarity(c + r, t) (heads) for every entity c,
    sorted descending, top_k.  Returns (predictions [n, top_k] int64, scores [n, top_k] fp32) like the reference's
    `.predictions` / `.scores`.  The all-candidates score matrix [n, n_ent] is never built."""
    if missing not in ("heads", "tails"):
        raise ValueError("missing entity should either be 'heads' or 'tails'")
    if dissimilarity_type not in ("L1", "L2"):
        raise ValueError("dissimilarity_type must be 'L1' or 'L2'")
    e = ent_emb[known_entities]
    r = rel_emb[known_relations]
    queries = e + r if missing == "tails" else e - r      # ||c + r - t|| = ||c - (t - r)||
    own = index is None
    index = CatalogIndex(ent_emb) if own else index
    try:
        dist, rows = index.topk_dissimilarity(queries, top_k, p=1 if dissimilarity_type == "L1" else 2)
    finally:
        if own:
            index.close()
    return rows, -dist
