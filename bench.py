#!/usr/bin/env python
"""Benchmark of the two-tower vector-similarity hot path (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-retrieval]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Headline line (one JSON object on stdout, rank 0):
  metric   pairs/s of the fused score+loss forward+backward, BASELINE config 2
           (two_tower inner-product + bce, 65 536 pairs x 1024-d bf16), inputs resident in HBM
  e2e      same metric through the HOST-buffer C-ABI entry point (pinned host inputs, H2D inside the timed
           region, loss read back)
  roofline achieved HBM GB/s of the fused kernel = algorithmic bytes (4*D*e + 16 per pair) / CUDA-event time
  cpu_baseline  the oracle port of the reference's modules timed on this box's host cores
  projection    SURVEY 8f rank 1 (head projection + score in one launch) with the library sequence beside it (N = 1 only)
  retrieval     BASELINE config 4 (cosine top-100, 1M x 1024 bf16 catalog, 10 000 queries), catalog rows sharded
                over the N ranks, per-shard top-k merged after one NCCL all-gather: queries/s + tensor roofline
With N > 1 the pair path is replicated (weak scaling, no data-path collective); retrieval is strong scaling.
`--impl reference` times the reference's own CPU implementation (oracle port: same torch ops, all host threads).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PAIRS, DIM = 65536, 1024                 # BASELINE config 2
CAT_ROWS, N_QUERIES, TOPK = 1_000_000, 10_000, 100   # BASELINE config 4
SEED = 20221009


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU work runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return None
        sm = []
        for r in rows:
            try:
                sm.append(float(r[0]))
            except ValueError:
                pass
        reasons = []
        for i, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in rows):
                reasons.append(name)
        try:
            mx = float(rows[0][1])
        except ValueError:
            mx = None
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(rows)}


def make_pairs(torch, device, seed, dtype):
    """SURVEY 8d recipe: x = tanh(z); positives y = tanh(z + 0.25 n), negatives independent; labels ~ Bernoulli(0.5)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    z = torch.randn(N_PAIRS, DIM, device=device, generator=gen)
    labels = (torch.rand(N_PAIRS, device=device, generator=gen) < 0.5).long()
    y = torch.where(labels[:, None] == 1, torch.tanh(z + 0.25 * torch.randn(N_PAIRS, DIM, device=device, generator=gen)),
                    torch.tanh(torch.randn(N_PAIRS, DIM, device=device, generator=gen))).to(dtype)
    x = torch.tanh(z).to(dtype)
    return x, y, labels


def cpu_reference_step(torch, x, y, labels):
    """The reference's CPU path for this step: head similarity + ladder + backward (oracle/torch_port.py restates
    src/models/base.py:29-34,77-86, text.py:1468-1477 from the same torch ops)."""
    from oracle import torch_port
    return torch_port.pair_score_loss_fwd_bwd("inner_product", "bce", x, y, labels)


def run_reference(args, rank, emit):
    import torch
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen = torch.Generator().manual_seed(SEED + 2000)
    rows = N_PAIRS
    x = torch.tanh(torch.randn(rows, DIM, generator=gen)).to(torch.bfloat16)
    y = torch.tanh(torch.randn(rows, DIM, generator=gen)).to(torch.bfloat16)
    labels = (torch.rand(rows, generator=gen) < 0.5).long()
    t = time.perf_counter(); cpu_reference_step(torch, x, y, labels); t1 = time.perf_counter() - t
    budget = 150.0
    total = args.steps + args.warmup
    if total * t1 > budget:
        rows = max(1024, int(rows * budget / (total * t1)) // 1024 * 1024)
        x, y, labels = x[:rows].contiguous(), y[:rows].contiguous(), labels[:rows].contiguous()
    for _ in range(args.warmup):
        cpu_reference_step(torch, x, y, labels)
    t = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(torch, x, y, labels)
    dt = time.perf_counter() - t
    val = rows * args.steps / dt
    sample = f"{rows} of {N_PAIRS} pairs x {DIM}-d per step, bf16 inputs upcast to fp32 as the reference's autocast does, {args.steps} steps"
    emit(({
        "impl": "reference", "metric": "pairs/s (fused score+loss fwd/bwd)", "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "two_tower inner_product + bce loss fwd/bwd, 65536 pairs x 1024-d bf16 (BASELINE config 2), CPU"},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def bench_retrieval(torch, dist, ia, device, rank, world, pk):
    """BASELINE config 4: cosine top-100 over a 1M x 1024 bf16 catalog, 10k queries; rows sharded over ranks."""
    lo, hi = ia.shard_bounds(CAT_ROWS, world, rank)
    gen = torch.Generator(device=device).manual_seed(SEED + 4000 + rank)
    cat = torch.empty((hi - lo, DIM), dtype=torch.bfloat16, device=device)
    for s in range(0, hi - lo, 131072):
        e = min(s + 131072, hi - lo)
        cat[s:e] = torch.tanh(torch.randn(e - s, DIM, device=device, generator=gen)).to(torch.bfloat16)
    if rank == 0:
        cat[5000:6000] = cat[:1000]          # exact duplicates: ties are part of the workload
    # queries: the same on every rank (seeded identically): rows of rank 0's shard + noise
    qgen = torch.Generator(device=device).manual_seed(SEED + 4999)
    base = torch.tanh(torch.randn(N_QUERIES, DIM, device=device, generator=qgen))
    queries = torch.tanh(base + 0.1 * torch.randn(N_QUERIES, DIM, device=device, generator=qgen)).to(torch.bfloat16)
    index = ia.ShardedCatalogIndex(cat, CAT_ROWS) if world > 1 else ia.CatalogIndex(cat)
    before = ia.launch_count()
    r_warm, r_steps = 3, 10
    for _ in range(r_warm):                                 # full-size warm-up passes (scratch growth, descriptors, NCCL buffers)
        index.topk_keys(queries, TOPK, "cosine")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ia.launch_count()
    e0.record()
    for _ in range(r_steps):
        keys = index.topk_keys(queries, TOPK, "cosine")
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / r_steps
    launches = ia.launch_count() - l0
    scores, rows = ia.unpack_keys(keys, "cosine")
    sorted_ok = bool((scores[:, :-1] >= scores[:, 1:]).all())
    index.close()
    flops = 2.0 * N_QUERIES * CAT_ROWS * DIM
    tf = flops / (ms * 1e-3) / 1e12 / world            # per-GPU achieved
    return {
        "metric": "retrieval queries/s @1M x 1024, top-100", "value": N_QUERIES / (ms * 1e-3), "unit": "queries/s",
        "ms_per_step": ms, "steps": r_steps, "warmup": r_warm, "scaling": "strong", "n_gpus": world, "dtype": "bf16",
        "config": {"workload": "cosine all-pairs same-item retrieval, 1M-item catalog x 10k queries x 1024-d bf16, top-100 "
                               "(BASELINE config 4)", "sharding": f"catalog rows over {world} rank(s), one NCCL all-gather of u64 keys",
                   "step": "one pass of all 10k queries over the whole catalog (all-pairs scores + top-100 + shard merge)"},
        "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tf / pk["tf_sustained"],
                     "traffic": None, "per_gpu": True, "peak_source": pk["source"] + " (cuBLAS bf16 sustained)",
                     "frac_of_burst": tf / pk["tf_burst"], "frac_of_nominal_2250": tf / 2250.0},
        "gpu_launches": launches, "sorted": sorted_ok,
    }


def bench_projection(torch, ia, device, pk):
    """head projection (dense -> tanh, both sides) + cosine score from the encoder features in one
    launch, 65 536 pairs, K = H = 1024, bf16; the library sequence the reference runs is timed beside it."""
    import item_alignment_b200.functional as F_
    n, k, h = N_PAIRS, DIM, DIM
    gen = torch.Generator(device=device).manual_seed(SEED + 6000)
    f1 = torch.randn(n, k, device=device, generator=gen).to(torch.bfloat16)
    f2 = torch.randn(n, k, device=device, generator=gen).to(torch.bfloat16)
    w = (torch.randn(h, k, device=device, generator=gen) / k ** 0.5).to(torch.bfloat16)
    b = torch.randn(h, device=device, generator=gen) * 0.1
    bt = b.to(torch.bfloat16)

    def timed(fn, warm=3, steps=20):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def library():
        x = torch.tanh(torch.nn.functional.linear(f1, w, bt))
        y = torch.tanh(torch.nn.functional.linear(f2, w, bt))
        s = torch.nn.functional.cosine_similarity(x, y)
        return s, (s + 1) / 2

    l0 = ia.launch_count()
    ms_score = timed(lambda: F_.project_score_raw("cosine", f1, f2, w, b))
    launches = ia.launch_count() - l0
    ms_proj = timed(lambda: F_.project_tanh_raw(f1, f2, w, b))
    ms_lib = timed(library)
    flops = 2.0 * 2 * n * k * h
    tf = flops / (ms_score * 1e-3) / 1e12
    return {
        "metric": "pairs/s (head projection + cosine score from encoder features)", "value": n / (ms_score * 1e-3), "unit": "pairs/s",
        "ms_per_step": ms_score, "steps": 20, "warmup": 3, "dtype": "bf16",
        "config": {"workload": "VecSimClassificationHead inference: tanh(dense(f)) for both sides + cosine score + probability, "
                               "65536 pairs, K = H = 1024 (SURVEY 8f rank 1); embeddings not written"},
        "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tf / pk["tf_sustained"],
                     "traffic": None, "peak_source": pk["source"] + " (cuBLAS bf16 sustained)", "frac_of_burst": tf / pk["tf_burst"]},
        "projection_only_ms": ms_proj, "library_linear_tanh_cosine_ms": ms_lib, "gpu_launches": launches,
    }


def main():
    # stdout carries exactly ONE JSON line: everything libraries print (e.g. "NCCL version ...") goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, emit)
        return

    import torch
    import torch.distributed as dist
    import item_alignment_b200 as ia
    from item_alignment_b200 import _lib
    from item_alignment_b200 import functional as F_

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    pk = peaks()
    lib = _lib.lib()
    t_start = time.time()
    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---------------------------------------------------------------- headline: fused score+loss fwd/bwd (config 2)
    x, y, labels = make_pairs(torch, device, SEED + 2000 + rank, torch.bfloat16)
    step = lambda: F_.pair_score_loss_raw("inner_product", "bce", x, y, labels, 1.0, "mean")   # one launch: loss + dx + dy
    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ia.launch_count()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = ia.launch_count() - l0
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total) / args.steps
    value = N_PAIRS * world / (ms_step * 1e-3)
    loss_val = float(out[2])
    alg_bytes = N_PAIRS * (4 * DIM * 2 + 16)                      # read x,y + write dx,dy (bf16) + label + sim,probs
    gbs = alg_bytes / (ms_step * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get("pair_fused_bf16_inner_bce_dram_bytes_per_launch")

    # ---------------------------------------------------------------- e2e: host buffers through the C ABI
    xh, yh, lh = x.cpu().pin_memory(), y.cpu().pin_memory(), labels.cpu().pin_memory()
    loss_h = torch.zeros(1).pin_memory()

    def e2e_step():
        _lib.check(lib.ia_pair_score_loss_host(0, 0, 1.0, 1, _lib.IA_BF16, xh.data_ptr(), yh.data_ptr(), lh.data_ptr(), N_PAIRS, DIM,
                                               loss_h.data_ptr(), None, None, local_rank))
        return float(loss_h[0])
    for _ in range(3):
        e2e_loss = e2e_step()
    e2e_steps = max(3, min(args.steps, 50))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_loss = e2e_step()            # returns after the loss is on the host (stream sync inside)
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], device=device)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = N_PAIRS * world / (float(e2e_ms) * 1e-3)
    n_chunks = -(-N_PAIRS // max(1024, (16 << 20) // (DIM * 2)))
    e2e = {"value": e2e_val, "unit": "pairs/s", "h2d_bytes_per_step": 2 * N_PAIRS * DIM * 2 + N_PAIRS * 8,
           "d2h_bytes_per_step": 4 * n_chunks, "ms_per_step": float(e2e_ms), "steps": e2e_steps,
           "api": "ia_pair_score_loss_host (pinned host x, y, labels -> loss on host)", "loss": e2e_loss}

    # ---------------------------------------------------------------- retrieval (config 4)
    retrieval = None
    if not args.no_retrieval:
        del x, y
        torch.cuda.empty_cache()
        retrieval = bench_retrieval(torch, dist, ia, device, rank, world, pk)
    projection = None
    if not args.no_retrieval and world == 1:
        torch.cuda.empty_cache()
        projection = bench_projection(torch, ia, device, pk)
    t_end = time.time()
    clocks = sampler.stop(t_start, t_end) if sampler else None

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        xc, yc, lc = xh.clone(), yh.clone(), lh.clone()
        cpu_reference_step(torch, xc, yc, lc)
        t0 = time.perf_counter(); reps = 0
        while True:
            cpu_reference_step(torch, xc, yc, lc); reps += 1
            if time.perf_counter() - t0 > 10.0 or reps >= 200:
                break
        dt = time.perf_counter() - t0
        cpu = {"value": N_PAIRS * reps / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": f"{reps} full steps of the same workload (65536 pairs x 1024-d) in {dt:.1f} s through oracle/torch_port.py "
                         "(the reference's torch ops, inputs upcast to fp32)"}

    if rank == 0:
        line = {
            "metric": "pairs/s (fused score+loss fwd/bwd)", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "two_tower inner_product + bce loss fwd/bwd, 65536 pairs x 1024-d bf16 (BASELINE config 2)",
                       "pairs_per_gpu": N_PAIRS, "dim": DIM, "l2": "working set 512 MiB per step (x, y, dx, dy) > 126 MB L2, no flush",
                       "parallelism": f"replicas x{world} (pairs independent, no collective)"},
            "loss": loss_val, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": pk["source"] + " (copy bandwidth)",
                         "frac_of_nominal_8000": gbs / 8000.0, "per_gpu": True},
            "cpu_baseline": cpu, "retrieval": retrieval, "projection": projection,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
