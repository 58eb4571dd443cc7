#!/usr/bin/env python
"""Softmax head (+ CE) timings, CUDA events: training and forward-only kernels for a few (dtype, h)."""
import sys

import torch

sys.path.insert(0, ".")
from item_alignment_b200 import functional as F_  # noqa: E402

dev = "cuda:0"
n = 65536


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def timed_graph(fn, iters=30, warm=3):
    """The same launches replayed from a CUDA graph: what the GPU needs, without the Python / allocator time of the eager call
    (the eager step of a 50 us kernel is bound by ~60 us of host work per call)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        keep = fn()
    for _ in range(warm):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    del keep
    return e0.elapsed_time(e1) / iters * 1e3


for dt, h in ((torch.bfloat16, 1024), (torch.bfloat16, 768), (torch.float32, 1024), (torch.float32, 768), (torch.float16, 512), (torch.bfloat16, 512)):
    x = torch.tanh(torch.randn(n, h, device=dev)).to(dt)
    y = torch.tanh(torch.randn(n, h, device=dev)).to(dt)
    l = (torch.rand(n, device=dev) < 0.5).long()
    w = torch.randn(2, 2 * h, device=dev) * 0.02
    b = torch.zeros(2, device=dev)
    e = x.element_size()
    t_train = timed(lambda: F_.softmax_head_raw(x, y, w, b, l))
    t_graph = timed_graph(lambda: F_.softmax_head_raw(x, y, w, b, l))
    t_main = timed(lambda: F_.softmax_head_raw(x, y, w, b, l, want_wgrads=False))
    t_fwd = timed(lambda: F_.softmax_head_raw(x, y, w, b))
    by_train, by_fwd = n * (4 * h * e + 24), n * (2 * h * e + 16)
    print(f"{str(dt):16s} h={h:5d}: train, graph replay {t_graph:7.1f} us ({by_train / t_graph / 1e3:6.0f} GB/s, {by_train / t_graph / 1e3 / 65.549:5.1f} %)  eager {t_train:7.1f} us "
          f"({by_train / t_train / 1e3 / 65.549:5.1f} %; without the dW finalize launch {t_main:6.1f} us)   "
          f"forward {t_fwd:6.1f} us ({by_fwd / t_fwd / 1e3:6.0f} GB/s, {by_fwd / t_fwd / 1e3 / 65.549:5.1f} %)", flush=True)
