#!/usr/bin/env python
"""Where the projection kernel's time goes: cycle counters of the MMA thread and the epilogue warps (IA_PROJ_DEBUG=2).
The epilogue-free comparison recorded in profiles/r01/projection_stats.log came from a temporary build that skipped
the epilogue arithmetic (186-194 us: the mainloop alone runs at the power-limited tensor rate)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import item_alignment_b200.functional as F_  # noqa: E402
from item_alignment_b200._lib import lib  # noqa: E402

n, k, h = 65536, 1024, 1024
dev = "cuda:0"
gen = torch.Generator(device=dev).manual_seed(1)
f1 = torch.randn(n, k, device=dev, generator=gen).bfloat16()
f2 = torch.randn(n, k, device=dev, generator=gen).bfloat16()
w = (torch.randn(h, k, device=dev, generator=gen) / k ** 0.5).bfloat16()
b = torch.randn(h, device=dev, generator=gen) * 0.1


def run(tag, fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    out = np.zeros(8, dtype=np.uint64)
    lib().ia_project_last_stats(out.ctypes.data_as(ctypes.c_void_p))
    s = out.astype(np.float64)
    msg = f"{tag}: {ms * 1e3:.1f} us"
    if s[2] > 0:
        msg += (f" | MMA thread: wait epilogue {100 * s[0] / s[2]:.1f}%, wait TMA {100 * s[1] / s[2]:.1f}% of {s[2] / 148:.0f} cycles/CTA"
                f" | epilogue warps: wait MMA {100 * s[3] / max(s[4], 1):.1f}% of {s[4] / 148 / 8:.0f} cycles/warp")
    print(msg, flush=True)


for flags in ("2",):
    os.environ["IA_PROJ_DEBUG"] = flags
    print("IA_PROJ_DEBUG =", flags)
    run("project_tanh", lambda: F_.project_tanh_raw(f1, f2, w, b))
    run("project_score cosine, no embeds", lambda: F_.project_score_raw("cosine", f1, f2, w, b))
    run("project_score inner, no embeds", lambda: F_.project_score_raw("inner_product", f1, f2, w, b))
    run("project_score cosine + embeds", lambda: F_.project_score_raw("cosine", f1, f2, w, b, want_embeds=True))
