"""CTA-pair (cta_group::2) form of the tensor-core retrieval kernel vs the single-CTA form: small-shape equality first (so that a
protocol bug shows up in seconds), then BASELINE config 4 timings.  IA_RETR_PAIR is read per call."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import item_alignment_b200 as ia

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(1)


def both(index, q, k, measure):
    out = {}
    for pair in ("0", "1"):
        os.environ["IA_RETR_PAIR"] = pair
        out[pair] = index.topk_keys(q, k, measure)
        torch.cuda.synchronize()
    return out


ok = True
for (C, Q, D, k, measure, dt) in ((4096, 256, 64, 10, "inner_product", torch.bfloat16), (5000, 300, 128, 10, "cosine", torch.bfloat16),
                                  (70000, 1000, 256, 100, "cosine", torch.float16), (100000, 129, 512, 32, "inner_product", torch.bfloat16),
                                  (300000, 2100, 1024, 100, "cosine", torch.bfloat16), (262144, 4096, 64, 128, "cosine", torch.bfloat16)):
    cat = torch.tanh(torch.randn(C, D, device=dev, generator=gen)).to(dt)
    q = torch.tanh(torch.randn(Q, D, device=dev, generator=gen)).to(dt)
    q[: Q // 4] = cat[: Q // 4]
    with ia.CatalogIndex(cat) as index:
        r = both(index, q, k, measure)
        same = bool(torch.equal(r["0"], r["1"]))
        ok &= same
        print(f"C={C} Q={Q} D={D} k={k} {measure} {dt}: pair == single-CTA keys: {same}", flush=True)
        if not same:
            bad = (r["0"] != r["1"]).any(dim=1).nonzero().flatten()
            print("   differing queries:", bad[:16].tolist(), "of", bad.numel())
if not ok:
    sys.exit(1)

C, Q, D = 1_000_000, 10_000, 1024
cat = torch.empty(C, D, dtype=torch.bfloat16, device=dev)
for s in range(0, C, 131072):
    cat[s:s + 131072] = torch.tanh(torch.randn(min(131072, C - s), D, device=dev, generator=gen)).to(torch.bfloat16)
q = torch.tanh(torch.randn(Q, D, device=dev, generator=gen)).to(torch.bfloat16)
q[:1000] = cat[:1000]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with ia.CatalogIndex(cat) as index:
    for k in (100, 10):
        res = {}
        for rep in range(2):
            for pair in ("0", "1"):
                os.environ["IA_RETR_PAIR"] = pair
                ts = []
                for it in range(8):
                    e0.record(); keys = index.topk_keys(q, k, "cosine"); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
                st = index.last_stats()
                ms = statistics.median(ts[2:])
                res[pair] = keys
                print(f"k={k:3d} pair={pair} {ms:7.3f} ms  {2.0 * Q * C * D / ms / 1e9:7.0f} TFLOP/s  appends/query {st['appends'] / Q:6.0f}  merges/query {st['compactions'] / Q:5.1f}  "
                      f"splits {st['splits']}x{st['tiles_per_split']}", flush=True)
        print("   identical keys:", bool(torch.equal(res["0"], res["1"])))

# a slice of BASELINE config 5 (100 000 queries x 512, top-10, inner product): 2 M of a shard's 12.5 M rows
del cat, q
C, Q, D = 2_000_000, 100_000, 512
cat = torch.empty(C, D, dtype=torch.bfloat16, device=dev)
for s in range(0, C, 131072):
    cat[s:s + 131072] = torch.tanh(torch.randn(min(131072, C - s), D, device=dev, generator=gen)).to(torch.bfloat16)
q = torch.tanh(torch.randn(Q, D, device=dev, generator=gen)).to(torch.bfloat16)
with ia.CatalogIndex(cat) as index:
    res = {}
    for rep in range(2):
        for pair in ("0", "1"):
            os.environ["IA_RETR_PAIR"] = pair
            ts = []
            for it in range(4):
                e0.record(); keys = index.topk_keys(q, 10, "inner_product"); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            st = index.last_stats()
            ms = statistics.median(ts[1:])
            res[pair] = keys
            print(f"C5 slice (2M x 512, 100k queries, k=10) pair={pair} {ms:8.3f} ms  {2.0 * Q * C * D / ms / 1e9:7.0f} TFLOP/s  splits {st['splits']}x{st['tiles_per_split']}", flush=True)
    print("   identical keys:", bool(torch.equal(res["0"], res["1"])))
