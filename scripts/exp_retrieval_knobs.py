"""C4 on one GPU (k = 100): sweep of the run-time knobs of the tensor-core retrieval path (IA_RETR_FLAGS bits, split-planner penalty).
IA_RETR_FLAGS is read per call; IA_RETR_PENALTY too."""
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
ref = None
with ia.CatalogIndex(cat) as index:
    for rep in range(2):
        for flags, pen in ((14, None), (30, None), (6, None), (10, None), (14, "0.02"), (14, "0.1"), (14, "0.035")):
            os.environ["IA_RETR_FLAGS"] = str(flags)
            if pen is None: os.environ.pop("IA_RETR_PENALTY", None)
            else: os.environ["IA_RETR_PENALTY"] = pen
            ts = []
            for it in range(7):
                e0.record(); keys = index.topk_keys(q, 100, "cosine"); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            st = index.last_stats()
            if ref is None: ref = keys
            print(f"flags={flags:3d} penalty={pen or 'default':7s} {statistics.median(ts[2:]):7.3f} ms  appends/query {st['appends'] / Q:5.0f} merges/query {st['compactions'] / Q:5.1f} "
                  f"splits {st['splits']}x{st['tiles_per_split']}  same keys {bool(torch.equal(keys, ref))}", flush=True)
