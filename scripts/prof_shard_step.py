"""One rank's work of an N-way row-sharded retrieval step, on one GPU, stage by stage (no NCCL)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import item_alignment_b200 as ia

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0 / 16
C, Q, D, K = 1_000_000, 10_000, 1024, 100
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(1)
rows = C // world
cat = torch.tanh(torch.randn(rows, D, device=dev, generator=gen)).to(torch.bfloat16)
q = torch.tanh(torch.randn(Q, D, device=dev, generator=gen)).to(torch.bfloat16)
local = ia.CatalogIndex(cat, row_base=0)
groups = max(1, -(-K // (16 * world)))
kp = -(-K // (world * groups))
per = max(2048, int(rows * frac) // groups) // 256 * 256
n_probe = per * groups
names = ["probe bound (one launch)", "-", "main topk (seeded)", "fake gather+merge"]
acc = {n: [] for n in names}
tot = []
for it in range(12):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    words = local.probe_bound(q, kp, groups, per, "cosine")
    ev[1].record()
    ev[2].record()
    keys = local.topk_keys(q, K, "cosine", init_tau=words)
    ev[3].record()
    parts = keys.unsqueeze(0).expand(world, Q, K).contiguous()
    merged = ia.merge_keys(parts, K)
    ev[4].record()
    torch.cuda.synchronize()
    if it >= 2:
        for i, n in enumerate(names):
            acc[n].append(ev[i].elapsed_time(ev[i + 1]))
        tot.append(ev[0].elapsed_time(ev[4]))
st = local.last_stats()
print(f"world={world} shard rows={rows} probe rows={n_probe} in {groups} group(s) k'={kp}")
for n in names:
    print(f"  {n:24s} {statistics.median(acc[n]):.3f} ms")
print(f"  total {statistics.median(tot):.3f} ms; main-pass appends/query {st['appends']/Q:.0f}, merges/query {st['compactions']/Q:.1f}, splits {st['splits']}x{st['tiles_per_split']}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for it in range(8):
    e0.record(); local.topk_keys_unprobed(q, K, "cosine"); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"  unseeded main pass {statistics.median(ts[2:]):.3f} ms; appends/query {local.last_stats()['appends']/Q:.0f}")
ts = []
for it in range(8):
    e0.record(); local.topk_keys(q, K, "cosine"); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"  self-probed pass (ia_catalog_topk default) {statistics.median(ts[2:]):.3f} ms; appends/query {local.last_stats()['appends']/Q:.0f}")
