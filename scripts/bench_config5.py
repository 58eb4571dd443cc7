#!/usr/bin/env python
"""BASELINE config 5: inner-product catalog retrieval, 100M-item catalog sharded by rows across 8xB200,
100k queries x 512-d bf16, top-10 with NCCL all-gather merge.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_config5.py [--rows-per-gpu N]

Each rank generates its 12.5M x 512 shard on the device from seed + rank (12.8 GB), queries are identical on all
ranks.  One untimed pass over an 8192-query slab + one full pass, then ONE timed full pass (all 100k queries against
all 100M rows, probe + scan + all-gather + merge).  A spot check recomputes 16 queries with torch matmuls per shard.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import item_alignment_b200 as ia

ap = argparse.ArgumentParser()
ap.add_argument("--rows-per-gpu", type=int, default=12_500_000)
ap.add_argument("--queries", type=int, default=100_000)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--k", type=int, default=10)
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
total = args.rows_per_gpu * world
lo, hi = ia.shard_bounds(total, world, rank)
gen = torch.Generator(device=dev).manual_seed(20221009 + 5000 + rank)
cat = torch.empty((hi - lo, args.dim), dtype=torch.bfloat16, device=dev)
for s in range(0, hi - lo, 1 << 20):
    e = min(s + (1 << 20), hi - lo)
    cat[s:e] = torch.tanh(torch.randn(e - s, args.dim, device=dev, generator=gen)).to(torch.bfloat16)
qgen = torch.Generator(device=dev).manual_seed(20221009 + 5999)
q = torch.tanh(torch.randn(args.queries, args.dim, device=dev, generator=qgen)).to(torch.bfloat16)
index = ia.ShardedCatalogIndex(cat, total) if world > 1 else ia.CatalogIndex(cat)
index.topk_keys(q[:8192], args.k, "inner_product")
index.topk_keys(q, args.k, "inner_product")
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
keys = index.topk_keys(q, args.k, "inner_product")
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms)
scores, rows = ia.unpack_keys(keys, "inner_product")
# spot check: 16 queries, exact per-shard top-k with torch, gathered and merged on the host
sub = q[:16].float()
best_s, best_i = [], []
for s in range(0, hi - lo, 1 << 20):
    e = min(s + (1 << 20), hi - lo)
    sc = sub @ cat[s:e].float().t()
    v, i = sc.topk(args.k, dim=1)
    best_s.append(v); best_i.append(i + lo + s)
v = torch.cat(best_s, 1); i = torch.cat(best_i, 1)
if world > 1:
    vs = [torch.empty_like(v) for _ in range(world)]; is_ = [torch.empty_like(i) for _ in range(world)]
    dist.all_gather(vs, v); dist.all_gather(is_, i)
    v = torch.cat(vs, 1); i = torch.cat(is_, 1)
order = torch.argsort(v, dim=1, descending=True, stable=True)[:, :args.k]
ref_s = torch.gather(v, 1, order); ref_i = torch.gather(i, 1, order)
ok_scores = bool(torch.allclose(scores[:16], ref_s, rtol=1e-5, atol=1e-3))
ok_rows = float((rows[:16] == ref_i).float().mean())
if rank == 0:
    flops = 2.0 * args.queries * total * args.dim
    print(json.dumps({"config": "BASELINE config 5 (inner product, top-%d, %d x %d-d bf16 catalog over %d GPU(s), %d queries)" % (args.k, total, args.dim, world, args.queries),
                      "ms_per_pass": ms, "queries_per_s": args.queries / ms * 1e3, "tflops_per_gpu": flops / ms / 1e9 / world,
                      "spot_check_scores_ok": ok_scores, "spot_check_row_match_fraction": ok_rows,
                      "sorted": bool((scores[:, :-1] >= scores[:, 1:]).all())}))
if world > 1: dist.destroy_process_group()
