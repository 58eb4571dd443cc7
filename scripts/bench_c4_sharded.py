#!/usr/bin/env python
"""BASELINE config 4 row-sharded over the ranks of a torchrun launch: A/B of the exchange strategies of ShardedCatalogIndex.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_c4_sharded.py"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import item_alignment_b200 as ia
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lo, hi = ia.shard_bounds(bench.CAT_ROWS, world, rank)
gen = torch.Generator(device=dev).manual_seed(bench.SEED + 4000 + rank)
cat = torch.empty((hi - lo, bench.DIM), dtype=torch.bfloat16, device=dev)
for s in range(0, hi - lo, 131072):
    e = min(s + 131072, hi - lo)
    cat[s:e] = torch.tanh(torch.randn(e - s, bench.DIM, device=dev, generator=gen)).to(torch.bfloat16)
qg = torch.Generator(device=dev).manual_seed(bench.SEED + 4999)
q = torch.tanh(torch.randn(bench.N_QUERIES, bench.DIM, device=dev, generator=qg)).to(torch.bfloat16)
index = ia.ShardedCatalogIndex(cat, bench.CAT_ROWS)
ref = None
for name, slice_merge, overlap, frac in (("all-gather, every rank merges all", False, False, 1 / 16), ("slice-wise merge (all-to-all + all-gather)", True, False, 1 / 16),
                                         ("slice-wise + overlapped halves", True, True, 1 / 16), ("slice-wise, probe 1/32", True, False, 1 / 32)):
    index.slice_merge, index.overlap, index.probe_fraction = slice_merge, overlap, frac
    for _ in range(3):
        keys = index.topk_keys(q, bench.TOPK, "cosine")
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        keys = index.topk_keys(q, bench.TOPK, "cosine")
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    same = True if ref is None else bool(torch.equal(keys, ref))
    ref = keys if ref is None else ref
    if rank == 0:
        print(f"N={world} {name:44s} {float(ms):.3f} ms/pass  {bench.N_QUERIES / float(ms) * 1e3 / 1e6:.3f} M q/s  same keys: {same}", flush=True)
index.close()
dist.destroy_process_group()
