"""Small retrieval workload for ncu captures (one tensor-core kernel launch of ~0.3 s)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import item_alignment_b200 as ia

q_n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
c_n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
d = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
measure = sys.argv[4] if len(sys.argv) > 4 else "cosine"
k = int(sys.argv[5]) if len(sys.argv) > 5 else 100
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 8
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(1)
cat = torch.tanh(torch.randn(c_n, d, device=dev, generator=gen)).to(torch.bfloat16)
q = torch.tanh(torch.randn(q_n, d, device=dev, generator=gen)).to(torch.bfloat16)
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
for _ in range(3): (a @ b)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): (a @ b)
e1.record(); torch.cuda.synchronize()
print(f"   calib: cuBLAS bf16 8192^3 {10*2*8192**3/e0.elapsed_time(e1)/1e9:.0f} TFLOP/s, flags={os.environ.get('IA_RETR_FLAGS','default')}")
del a, b
with ia.CatalogIndex(cat) as index:
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        keys = index.topk_keys(q, k, measure)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    st = index.last_stats()
    import statistics
    ms = statistics.median(times[1:])
    print(f"   stats per query: appends {st['appends']/q_n:.0f} compactions {st['compactions']/q_n:.1f}; rare groups {st['rare_groups']} iters {st['rare_blocks']}; splits {st['splits']} x {st['tiles_per_split']} tiles; min {min(times):.2f} ms; epilogue cycles: wait {100*st['cyc_wait']/max(1,st['cyc_total']):.0f}% merge {100*st['cyc_compact']/max(1,st['cyc_total']):.0f}% rare {100*st['cyc_rare']/max(1,st['cyc_total']):.0f}% total/warp {st['cyc_total']/592/1e6:.1f}M")
    if True:
        print(f"{measure} k={k} Q={q_n} C={c_n} D={d}: {ms:.2f} ms, {2.0*q_n*c_n*d/ms/1e9:.1f} TFLOP/s, {q_n/ms*1e3:.0f} q/s")
