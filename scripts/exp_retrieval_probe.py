"""BASELINE config 4 on one GPU: ia_catalog_topk with its own probe pass (default) vs the plain cold-start scan."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import item_alignment_b200 as ia
C, Q, D = 1_000_000, 10_000, 1024
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(1)
cat = torch.empty(C, D, dtype=torch.bfloat16, device=dev)
for s in range(0, C, 131072):
    cat[s:s + 131072] = torch.tanh(torch.randn(min(131072, C - s), D, device=dev, generator=gen)).to(torch.bfloat16)
q = torch.tanh(torch.randn(Q, D, device=dev, generator=gen)).to(torch.bfloat16)
q[:1000] = cat[:1000]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with ia.CatalogIndex(cat) as index:
    for k in (100, 128, 32, 10):
        res = {}
        for name, fn in (("probed (default)", index.topk_keys), ("plain", index.topk_keys_unprobed)):
            ts = []
            for it in range(8):
                e0.record(); keys = fn(q, k, "cosine"); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            st = index.last_stats()
            ms = statistics.median(ts[2:])
            res[name] = keys
            print(f"k={k:3d} {name:18s} {ms:7.3f} ms  {2.0 * Q * C * D / ms / 1e9:7.0f} TFLOP/s  appends/query {st['appends'] / Q:6.0f}  merges/query {st['compactions'] / Q:5.1f}  "
                  f"splits {st['splits']}x{st['tiles_per_split']}", flush=True)
        print("   identical keys:", bool(torch.equal(res["probed (default)"], res["plain"])))
