"""Where do the ~9 us between ncu's kernel time and the event-timed step go?  Eager launches vs CUDA-graph replay,
and batch-size scaling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import item_alignment_b200 as ia
from item_alignment_b200 import functional as F_

dev = torch.device("cuda:0")
for n in (65536, 262144):
    d = 1024
    x = torch.tanh(torch.randn(n, d, device=dev)).to(torch.bfloat16)
    y = torch.tanh(torch.randn(n, d, device=dev)).to(torch.bfloat16)
    labels = (torch.rand(n, device=dev) < 0.5).long()
    step = lambda: F_.pair_score_loss_raw("inner_product", "bce", x, y, labels, 1.0, "mean")
    for _ in range(20): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 300
    e0.record()
    for _ in range(K): step()
    e1.record(); torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / K * 1e3
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
        with torch.cuda.graph(g, stream=s):
            out = step()
    torch.cuda.synchronize()
    for _ in range(20): g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K): g.replay()
    e1.record(); torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / K * 1e3
    alg = n * (4 * d * 2 + 16)
    print(f"n={n}: eager {eager:.1f} us/step ({alg/eager/1e3:.0f} GB/s)  graph {graph:.1f} us/step ({alg/graph/1e3:.0f} GB/s)")
