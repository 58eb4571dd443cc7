#!/usr/bin/env python
"""Head projection fused in front of the score, timed with CUDA events on one B200:
  ours   ia_project_tanh_fwd (one tcgen05 launch for both sides) and ia_project_score_fwd (+ score, embeddings unwritten)
  lib    what the reference runs: F.linear (cuBLAS) + torch.tanh per side, then the pair score
Usage: python scripts/bench_projection.py [n] [k] [h] [dtype]"""
import json
import sys

import torch

sys.path.insert(0, ".")
import item_alignment_b200.functional as F_  # noqa: E402


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    dt = dict(bf16=torch.bfloat16, f16=torch.float16)[sys.argv[4] if len(sys.argv) > 4 else "bf16"]
    dev = "cuda:0"
    gen = torch.Generator(device=dev).manual_seed(1)
    f1 = torch.randn(n, k, device=dev, generator=gen).to(dt)
    f2 = torch.randn(n, k, device=dev, generator=gen).to(dt)
    w = (torch.randn(h, k, device=dev, generator=gen) / k ** 0.5).to(dt)
    b = torch.randn(h, device=dev, generator=gen) * 0.1
    bt = b.to(dt)
    flops = 2.0 * 2 * n * k * h
    out = {"n": n, "k": k, "h": h, "dtype": str(dt)}

    t = timed(lambda: F_.project_tanh_raw(f1, f2, w, b))
    out["ours_project_tanh_ms"] = t
    out["ours_project_tanh_tflops"] = flops / t / 1e9
    t = timed(lambda: F_.project_score_raw("cosine", f1, f2, w, b))
    out["ours_project_score_ms"] = t
    out["ours_project_score_tflops"] = flops / t / 1e9
    t = timed(lambda: F_.project_score_raw("cosine", f1, f2, w, b, want_embeds=True))
    out["ours_project_score_embeds_ms"] = t
    out["ours_project_tanh_fast_ms"] = timed(lambda: F_.project_tanh_raw(f1, f2, w, b, fast_tanh=True))
    out["ours_project_score_fast_ms"] = timed(lambda: F_.project_score_raw("cosine", f1, f2, w, b, fast_tanh=True))

    def lib_project():
        return torch.tanh(torch.nn.functional.linear(f1, w, bt)), torch.tanh(torch.nn.functional.linear(f2, w, bt))

    t = timed(lib_project)
    out["lib_linear_tanh_ms"] = t
    out["lib_linear_tanh_tflops"] = flops / t / 1e9
    t = timed(lambda: (torch.nn.functional.linear(f1, w, bt), torch.nn.functional.linear(f2, w, bt)))
    out["lib_linear_only_ms"] = t

    def lib_project_score():
        x, y = lib_project()
        s = torch.nn.functional.cosine_similarity(x, y)
        return s, (s + 1) / 2

    out["lib_project_score_ms"] = timed(lib_project_score)

    def ours_then_pair():
        x, y = F_.project_tanh_raw(f1, f2, w, b)
        return F_.pair_score_raw("cosine", x, y)

    out["ours_project_then_pair_kernel_ms"] = timed(ours_then_pair)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
