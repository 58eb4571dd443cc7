#!/usr/bin/env python
"""One launch of each projection kernel variant for ncu (scripts/bench_projection.py holds the timings)."""
import sys

import torch

sys.path.insert(0, ".")
import item_alignment_b200.functional as F_  # noqa: E402

n, k, h = 65536, 1024, 1024
dev = "cuda:0"
gen = torch.Generator(device=dev).manual_seed(1)
f1 = torch.randn(n, k, device=dev, generator=gen).bfloat16()
f2 = torch.randn(n, k, device=dev, generator=gen).bfloat16()
w = (torch.randn(h, k, device=dev, generator=gen) / k ** 0.5).bfloat16()
b = torch.randn(h, device=dev, generator=gen) * 0.1
for _ in range(2):
    F_.project_tanh_raw(f1, f2, w, b)
    F_.project_score_raw("cosine", f1, f2, w, b)
torch.cuda.synchronize()
