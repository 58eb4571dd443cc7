"""Tail quantisation of the fused pair kernel: ns per pair at batch sizes around a whole number of iterations per warp
(296 CTAs x 8 warps x 2 rows = 4736 pairs per grid sweep; 65 536 pairs = 13.84 sweeps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from item_alignment_b200 import functional as F_
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
sweep = 296 * 8 * 2
for n in (65536, 13 * sweep, 14 * sweep, 14 * sweep + 512, 66304 - 2368):
    x = torch.tanh(torch.randn(n, 1024, device=dev, generator=g)).bfloat16(); y = torch.tanh(torch.randn(n, 1024, device=dev, generator=g)).bfloat16()
    l = (torch.rand(n, device=dev, generator=g) < 0.5).long()
    fn = lambda: F_.pair_score_loss_raw("inner_product", "bce", x, y, l)
    for _ in range(5): fn()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        for _ in range(20): keep = fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gr.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 100 * 1e3
    print(f"n = {n:6d} ({n / sweep:6.3f} sweeps): {us:7.2f} us  {us / n * 1e3:6.4f} ns/pair  {n * 8208 / us / 1e3:6.0f} GB/s", flush=True)
