#!/usr/bin/env python
"""Benchmark of the two-tower vector-similarity hot path (contract: see the task statement / DESIGN.md section 7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--only-headline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

ONE JSON line on stdout (rank 0).  Top-level fields = BASELINE config 2, the headline:
  metric/value  pairs/s of the fused score+loss forward+backward (two_tower inner-product + bce, 65 536 pairs x 1024-d bf16),
                inputs resident in HBM, ONE kernel launch per step (sim, probs, loss, dx, dy)
                The K timed steps are captured into one CUDA graph and replayed (`step_launch`); the same K steps launched
                eagerly from Python are timed right after and reported as `eager_ms_per_step`
  e2e           the same step through the HOST-buffer C-ABI entry point ia_pair_score_loss_host: pinned host x, y, labels in,
                loss AND dx, dy back in pinned host memory -- H2D and D2H inside the timed region
  roofline      achieved HBM GB/s of the fused kernel = algorithmic bytes (4*D*e + 16 per pair) / CUDA-event time per launch
  cpu_baseline  the oracle port of the reference's modules (same torch ops) timed on this box's host cores (N = 1)
Every other BASELINE config is measured in the same run and summarised in `summary` -- the LAST key of the line:
  c1  cosine score + 0.5 threshold, 50 000 x 768 fp32 (forward kernel; also through ia_pair_score_host)
  c3  {l1, l2} x {hinge, euclidean} fused fwd/bwd, 65 536 x 1024 fp32
  c4  cosine top-100 retrieval, 1M x 1024 bf16 catalog, 10 000 queries, catalog rows sharded over the N ranks (strong scaling),
      with the sharded result compared bit for bit with a single-index pass (parity_ok) and with a torch fp32 oracle slab
  c5  inner-product top-10, 12.5M x 512 bf16 rows PER GPU (N = 8: the full 100M-row BASELINE config 5), 100 000 queries
  softmax_ce / projection  the "softmax" measure + CE training kernel and the fused head projection
With N > 1 the pair path is replicated (weak scaling, no data-path collective); retrieval shards rows + one NCCL all-gather.
`--impl reference` times the reference's own CPU implementation of config 2 (oracle port: the reference's torch ops, all host
threads) on the same input recipe.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PAIRS, DIM = 65536, 1024                 # BASELINE config 2 / 3
C1_PAIRS, C1_DIM = 50000, 768              # BASELINE config 1
CAT_ROWS, N_QUERIES, TOPK = 1_000_000, 10_000, 100   # BASELINE config 4
C5_ROWS_PER_GPU, C5_QUERIES, C5_DIM, C5_K = 12_500_000, 100_000, 512, 10   # BASELINE config 5 (per-GPU shard)
SEED = 20221009
WORKLOAD = "two_tower inner_product + bce loss fwd/bwd, 65536 pairs x 1024-d bf16 (BASELINE config 2)"


def config_of(world):
    """Identical in the native and the reference arm (the driver compares them)."""
    return {"workload": WORKLOAD, "pairs_per_gpu": N_PAIRS, "dim": DIM,
            "inputs": "x = tanh(z); positives y = tanh(z + 0.25 n), negatives independent; labels ~ Bernoulli(0.5); bf16",
            "l2": "working set 512 MiB per step (x, y, dx, dy) > 126 MB L2, no flush",
            "parallelism": f"replicas x{world} (pairs independent, no collective)"}


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


def make_pairs(torch, device, seed, dtype, n=N_PAIRS, d=DIM):
    """SURVEY 8d recipe: x = tanh(z); positives y = tanh(z + 0.25 n), negatives independent; labels ~ Bernoulli(0.5)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    z = torch.randn(n, d, device=device, generator=gen)
    labels = (torch.rand(n, device=device, generator=gen) < 0.5).long()
    y = torch.where(labels[:, None] == 1, torch.tanh(z + 0.25 * torch.randn(n, d, device=device, generator=gen)),
                    torch.tanh(torch.randn(n, d, device=device, generator=gen))).to(dtype)
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
    rows = N_PAIRS
    x, y, labels = make_pairs(torch, torch.device("cpu"), SEED + 2000, torch.bfloat16)      # the native arm's recipe
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
    sample = (f"{rows} of {N_PAIRS} pairs x {DIM}-d per step, bf16 inputs upcast to fp32 as the reference's autocast does, "
              f"forward + loss + backward, {args.steps} steps, {cores} torch threads")
    emit(({
        "impl": "reference", "metric": "pairs/s (fused score+loss fwd/bwd)", "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config_of(args.gpus),
        "compute": "CPU, fp32 arithmetic on the bf16-quantised inputs (oracle/torch_port.py: the reference's torch ops)",
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ helpers
def timed_us(torch, fn, iters, warm):
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


def timed_graph_us(torch, fn, iters, warm):
    """fn's launches replayed from a CUDA graph: GPU time of a step whose eager call is bound by host work (allocations, ctypes)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        keep = fn()
    us = timed_us(torch, graph.replay, iters, warm)
    del keep
    return us


def r3(v):
    return None if v is None else float(f"{v:.4g}")


def gpu_numa_node(torch, index):
    """NUMA node of the GPU's PCIe root (sysfs), or None."""
    try:
        p = torch.cuda.get_device_properties(index)
        path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


class NumaPreferred:
    """Allocate (and first-touch) host memory on the GPU's NUMA node: set_mempolicy(MPOL_PREFERRED, node) around the
    pinned allocations of the e2e leg -- every rank otherwise takes its pages from the node its launcher happens to run on,
    and at N = 8 all H2D/D2H traffic then funnels through one socket's memory controllers.  Best effort: a container
    without the syscall (or a single-node box) leaves the default policy in place and says so."""

    def __init__(self, node):
        self.node, self.state = node, "default"

    def __enter__(self):
        if self.node is None or not os.path.isdir(f"/sys/devices/system/node/node{self.node}"):
            return self
        try:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << self.node)
            rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))      # set_mempolicy(MPOL_PREFERRED)
            self.state = "preferred" if rc == 0 else f"set_mempolicy errno {ctypes.get_errno()}"
        except Exception as e:
            self.state = f"unavailable ({type(e).__name__})"
        return self

    def __exit__(self, *a):
        if self.state == "preferred":
            try:
                ctypes.CDLL(None).syscall(238, 0, None, ctypes.c_ulong(0))         # MPOL_DEFAULT
            except Exception:
                pass


def bench_pair_configs(torch, F_, _lib, device, local_rank, pk):
    """BASELINE configs 1 and 3 (device-resident, CUDA events) + the softmax-measure CE step."""
    lib = _lib.lib()
    out = {}
    # ---- C1: cosine score + 0.5 threshold labels, 50 000 x 768 fp32, forward kernel
    x, y, _ = make_pairs(torch, device, SEED + 1000, torch.float32, C1_PAIRS, C1_DIM)
    us = timed_us(torch, lambda: F_.pair_score_raw("cosine", x, y, threshold=0.5), 100, 10)
    b = C1_PAIRS * (2 * C1_DIM * 4 + 9)
    sim, probs, lab = F_.pair_score_raw("cosine", x, y, threshold=0.5)
    # the CPU-side caller's route: host buffers through ia_pair_score_host (H2D + kernel + D2H of sim, probs, labels)
    xh, yh = x.cpu().pin_memory(), y.cpu().pin_memory()
    sh, ph = torch.empty(C1_PAIRS).pin_memory(), torch.empty(C1_PAIRS).pin_memory()
    lh = torch.empty(C1_PAIRS, dtype=torch.uint8).pin_memory()

    def host_call():
        _lib.check(lib.ia_pair_score_host(1, _lib.IA_F32, xh.data_ptr(), yh.data_ptr(), C1_PAIRS, C1_DIM, sh.data_ptr(), ph.data_ptr(),
                                          0.5, lh.data_ptr(), local_rank))
    for _ in range(3):
        host_call()
    t0 = time.perf_counter()
    for _ in range(10):
        host_call()
    e2e_ms = (time.perf_counter() - t0) * 1e2
    same = bool(torch.equal(sh, sim.cpu()) and torch.equal(lh.bool(), lab.cpu()))
    out["c1"] = {"Mpairs_s": r3(C1_PAIRS / us), "us": r3(us), "hbm_GBs": r3(b / us / 1e3), "hbm_frac": r3(b / us / 1e3 / pk["hbm"]),
                 "e2e_host_Mpairs_s": r3(C1_PAIRS / e2e_ms / 1e3), "host_equals_device": same,
                 "positive_labels": int(lab.sum())}
    del x, y, xh, yh
    # ---- C3: {l1, l2} x {hinge, euclidean}, 65 536 x 1024 fp32, fused forward + backward
    x, y, labels = make_pairs(torch, device, SEED + 3000, torch.float32)
    b = N_PAIRS * (4 * DIM * 4 + 16)
    c3 = {}
    for m in ("l1", "l2"):
        for lo in ("hinge", "euclidean"):
            us = timed_us(torch, lambda: F_.pair_score_loss_raw(m, lo, x, y, labels), 50, 5)
            c3[f"{m}_{lo}"] = {"Mpairs_s": r3(N_PAIRS / us), "us": r3(us), "hbm_frac": r3(b / us / 1e3 / pk["hbm"])}
    out["c3"] = c3
    del x, y
    torch.cuda.empty_cache()
    # ---- "softmax" measure: TwoTowerClassificationHead + CrossEntropyLoss forward + backward (bf16)
    sm = {}
    for h in (1024, 768):
        x, y, labels = make_pairs(torch, device, SEED + 7000 + h, torch.bfloat16, N_PAIRS, h)
        w = torch.randn(2, 2 * h, device=device) * 0.02
        bb = torch.zeros(2, device=device)
        us_eager = timed_us(torch, lambda: F_.softmax_head_raw(x, y, w, bb, labels), 50, 5)
        us = us_eager
        if int(os.environ.get("WORLD_SIZE", "1")) == 1:
            # both launches (main + dW finalize) replayed from a CUDA graph: the eager call of a < 100 us step is host bound.
            # Single process only: stream capture and a live NCCL watchdog thread do not mix reliably.
            try:
                us = timed_graph_us(torch, lambda: F_.softmax_head_raw(x, y, w, bb, labels), 50, 5)
            except Exception:
                us = us_eager
        b = N_PAIRS * (4 * h * 2 + 24)
        sm[f"bf16_{h}"] = {"Mpairs_s": r3(N_PAIRS / us), "us": r3(us), "hbm_frac": r3(b / us / 1e3 / pk["hbm"]), "eager_us": r3(us_eager)}
        del x, y
    out["softmax_ce"] = sm
    torch.cuda.empty_cache()
    return out


def oracle_slab_check(torch, queries, cat_chunks, row_bases, keys, ia, k, cosine, n_check=32):
    """Independent check of a slab of retrieval results against torch fp32 matmuls + a stable sort (rank-local shards are
    passed as chunks with their global row bases).  Scores must agree to fp32 summation noise; rows must be identical
    wherever the oracle's neighbouring ranks are separated by more than that noise."""
    q = queries[:n_check].float()
    if cosine:
        q = q / q.norm(dim=1, keepdim=True).clamp_min(1e-8)
    best_s, best_i = [], []
    for cat, base in zip(cat_chunks, row_bases):
        for s in range(0, cat.shape[0], 262144):
            c = cat[s:s + 262144].float()
            if cosine:
                c = c / c.norm(dim=1, keepdim=True).clamp_min(1e-8)
            sc = q @ c.t()
            kk = min(k + 1, sc.shape[1])
            v, i = torch.sort(sc, dim=1, descending=True, stable=True)
            best_s.append(v[:, :kk]); best_i.append(i[:, :kk] + base + s)
    return torch.cat(best_s, 1), torch.cat(best_i, 1)


def finish_oracle_check(torch, cand_s, cand_i, keys, ia, k, measure, scale):
    """cand_*: gathered per-chunk candidates [n, *]; compares with our keys[:n]."""
    n = cand_s.shape[0]
    # global stable order: score descending, then row ascending
    order = torch.argsort(cand_i, dim=1, stable=True)
    cs, ci = torch.gather(cand_s, 1, order), torch.gather(cand_i, 1, order)
    order = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :k + 1]
    rs, ri = torch.gather(cs, 1, order), torch.gather(ci, 1, order)
    scores, rows = ia.unpack_keys(keys[:n].contiguous(), measure)
    tol = 4e-6 * scale
    scores_ok = bool(((scores - rs[:, :k]).abs() <= tol).all())
    gap_next = (rs[:, :k] - rs[:, 1:k + 1]).abs()
    gap_prev = torch.cat((torch.full_like(gap_next[:, :1], 1e9), gap_next[:, :-1]), 1)
    clear = (gap_next > 4 * tol) & (gap_prev > 4 * tol)
    rows_ok = bool((rows[clear] == ri[:, :k][clear]).all())
    return {"scores_ok": scores_ok, "rows_ok_where_separated": rows_ok, "rows_equal_frac": r3(float((rows == ri[:, :k]).float().mean())),
            "separated_frac": r3(float(clear.float().mean())), "queries": n}


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
    # queries: the same on every rank (seeded identically)
    qgen = torch.Generator(device=device).manual_seed(SEED + 4999)
    base = torch.tanh(torch.randn(N_QUERIES, DIM, device=device, generator=qgen))
    queries = torch.tanh(base + 0.1 * torch.randn(N_QUERIES, DIM, device=device, generator=qgen)).to(torch.bfloat16)
    queries[:1000] = cat[:1000] if rank == 0 else queries[:1000]     # rank 0's rows as queries: true neighbours + exact ties
    if world > 1:
        dist.broadcast(queries, 0)
    index = ia.ShardedCatalogIndex(cat, CAT_ROWS) if world > 1 else ia.CatalogIndex(cat)
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
    self_ok = bool((rows[:1000, 0] == torch.arange(1000, device=device)).all())     # a catalog row's best match is itself (lowest id of the tie)
    # ---- parity of the sharded result (outside the timed region)
    n_slab = 256
    parity_ok, parity_how = None, None
    cand_s, cand_i = oracle_slab_check(torch, queries, [cat], [lo], keys, ia, TOPK, True)
    if world > 1:
        # (a) bit-exact: every rank all-gathers the shards (2 GB) and rank 0 runs ONE single-index pass over the whole catalog
        full = torch.empty((CAT_ROWS, DIM), dtype=torch.bfloat16, device=device)
        per = -(-CAT_ROWS // world)
        padded = torch.zeros((per, DIM), dtype=torch.bfloat16, device=device)
        padded[:hi - lo] = cat
        gathered = torch.empty((world * per, DIM), dtype=torch.bfloat16, device=device)
        dist.all_gather_into_tensor(gathered, padded)
        for r in range(world):
            a, b = ia.shard_bounds(CAT_ROWS, world, r)
            full[a:b] = gathered[r * per:r * per + (b - a)]
        del gathered, padded
        if rank == 0:
            with ia.CatalogIndex(full) as single:
                ref_keys = single.topk_keys(queries[:n_slab], TOPK, "cosine")
            parity_ok = bool(torch.equal(ref_keys, keys[:n_slab]))
        parity_how = f"sharded+NCCL-merged keys of {n_slab} queries == single-index pass over the all-gathered catalog, bit for bit"
        del full
        # (b) oracle candidates of every shard -> rank 0
        gs = [torch.empty_like(cand_s) for _ in range(world)]
        gi = [torch.empty_like(cand_i) for _ in range(world)]
        dist.all_gather(gs, cand_s); dist.all_gather(gi, cand_i)
        cand_s, cand_i = torch.cat(gs, 1), torch.cat(gi, 1)
    else:
        # single GPU: seeded / split plan vs an unseeded pass of a different decomposition (query slab alone)
        ref_keys = index.topk_keys(queries[:n_slab], TOPK, "cosine")
        parity_ok = bool(torch.equal(ref_keys, keys[:n_slab]))
        parity_how = f"keys of {n_slab} queries from the 10k-query pass == a {n_slab}-query pass (different split plan), bit for bit"
    oracle = finish_oracle_check(torch, cand_s, cand_i, keys, ia, TOPK, "cosine", 1.0)
    index.close()
    flops = 2.0 * N_QUERIES * CAT_ROWS * DIM
    tf = flops / (ms * 1e-3) / 1e12 / world            # per-GPU achieved
    return {
        "metric": "retrieval queries/s @1M x 1024, top-100", "value": N_QUERIES / (ms * 1e-3), "unit": "queries/s",
        "ms_per_step": ms, "steps": r_steps, "warmup": r_warm, "scaling": "strong", "n_gpus": world, "dtype": "bf16",
        "config": {"workload": "cosine all-pairs same-item retrieval, 1M x 1024 bf16 catalog, 10k queries, top-100 (BASELINE config 4)",
                   "sharding": f"rows over {world} rank(s), one NCCL all-gather of u64 keys"},
        "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tf / pk["tf_sustained"],
                     "per_gpu": True, "peak_source": pk["source"] + " (cuBLAS bf16 sustained)",
                     "frac_of_burst": tf / pk["tf_burst"], "frac_of_nominal_2250": tf / 2250.0},
        "gpu_launches": launches, "sorted": sorted_ok, "self_match": self_ok, "parity_ok": parity_ok, "parity": parity_how,
        "oracle_slab": oracle,
    }


def bench_config5(torch, dist, ia, device, rank, world, pk):
    """BASELINE config 5: inner product, top-10, 512-d bf16, 100k queries; 12.5M catalog rows per GPU (N = 8 is the full
    100M-row catalog), row-sharded, per-shard top-k merged after one NCCL all-gather."""
    total = C5_ROWS_PER_GPU * world
    lo, hi = ia.shard_bounds(total, world, rank)
    gen = torch.Generator(device=device).manual_seed(SEED + 5000 + rank)
    cat = torch.empty((hi - lo, C5_DIM), dtype=torch.bfloat16, device=device)
    for s in range(0, hi - lo, 1 << 20):
        e = min(s + (1 << 20), hi - lo)
        cat[s:e] = torch.tanh(torch.randn(e - s, C5_DIM, device=device, generator=gen)).to(torch.bfloat16)
    qgen = torch.Generator(device=device).manual_seed(SEED + 5999)
    q = torch.tanh(torch.randn(C5_QUERIES, C5_DIM, device=device, generator=qgen)).to(torch.bfloat16)
    index = ia.ShardedCatalogIndex(cat, total) if world > 1 else ia.CatalogIndex(cat)
    index.topk_keys(q[:8192], C5_K, "inner_product")
    index.topk_keys(q, C5_K, "inner_product")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    steps = 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        keys = index.topk_keys(q, C5_K, "inner_product")
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / steps
    scores, _ = ia.unpack_keys(keys, "inner_product")
    sorted_ok = bool((scores[:, :-1] >= scores[:, 1:]).all())
    cand_s, cand_i = oracle_slab_check(torch, q, [cat], [lo], keys, ia, C5_K, False, n_check=16)
    if world > 1:
        gs = [torch.empty_like(cand_s) for _ in range(world)]
        gi = [torch.empty_like(cand_i) for _ in range(world)]
        dist.all_gather(gs, cand_s); dist.all_gather(gi, cand_i)
        cand_s, cand_i = torch.cat(gs, 1), torch.cat(gi, 1)
    oracle = finish_oracle_check(torch, cand_s, cand_i, keys, ia, C5_K, "inner_product", float(C5_DIM) ** 0.5 * 8)
    index.close()
    del cat
    torch.cuda.empty_cache()
    tf = 2.0 * C5_QUERIES * total * C5_DIM / (ms * 1e-3) / 1e12 / world
    return {"rows_total": total, "rows_per_gpu": C5_ROWS_PER_GPU, "queries": C5_QUERIES, "k": C5_K, "dim": C5_DIM, "n_gpus": world,
            "ms_per_pass": r3(ms), "qps": r3(C5_QUERIES / (ms * 1e-3)), "tf_per_gpu": r3(tf), "frac_sustained": r3(tf / pk["tf_sustained"]),
            "sorted": sorted_ok, "oracle_slab": oracle, "full_config5": world == 8}


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

    def library():
        x = torch.tanh(torch.nn.functional.linear(f1, w, bt))
        y = torch.tanh(torch.nn.functional.linear(f2, w, bt))
        s = torch.nn.functional.cosine_similarity(x, y)
        return s, (s + 1) / 2

    ms_score = timed_us(torch, lambda: F_.project_score_raw("cosine", f1, f2, w, b), 20, 3) / 1e3
    ms_proj = timed_us(torch, lambda: F_.project_tanh_raw(f1, f2, w, b), 20, 3) / 1e3
    ms_lib = timed_us(torch, library, 20, 3) / 1e3
    tf = 2.0 * 2 * n * k * h / (ms_score * 1e-3) / 1e12
    # training step of the head (train() mode, dropout 0.1): projection + cosine score + bce + the whole backward
    labels = (torch.rand(n, device=device, generator=gen) < 0.5).long()
    g1, g2 = f1.clone().requires_grad_(True), f2.clone().requires_grad_(True)
    wm, bm = w.float().requires_grad_(True), b.clone().requires_grad_(True)      # fp32 master weights, cast per step like autocast
    step_no = [0]

    def ours_train():
        step_no[0] += 1
        _, _, _, _, loss = F_.project_score_loss_train("cosine", "bce", g1, g2, wm, bm, labels, 0.1, SEED, step_no[0])
        loss.backward()
        g1.grad = g2.grad = wm.grad = bm.grad = None

    def library_train():
        drop = torch.nn.functional.dropout
        w16, b16 = wm.to(torch.bfloat16), bm.to(torch.bfloat16)
        x = drop(torch.tanh(torch.nn.functional.linear(drop(g1, 0.1), w16, b16)), 0.1)
        y = drop(torch.tanh(torch.nn.functional.linear(drop(g2, 0.1), w16, b16)), 0.1)
        s = torch.nn.functional.cosine_similarity(x.float(), y.float())
        torch.nn.functional.binary_cross_entropy_with_logits(s, labels.float()).backward()
        g1.grad = g2.grad = wm.grad = bm.grad = None

    ms_train = timed_us(torch, ours_train, 10, 3) / 1e3
    ms_train_lib = timed_us(torch, library_train, 10, 3) / 1e3
    tf_train = 3 * 2.0 * 2 * n * k * h / (ms_train * 1e-3) / 1e12          # forward + data-gradient + weight-gradient GEMMs
    return {"workload": "tanh(dense(f)) both sides + cosine score + probability in one launch, 65536 pairs, K = H = 1024 bf16",
            "Mpairs_s": r3(n / ms_score / 1e3), "ms": r3(ms_score), "tf": r3(tf), "frac_sustained": r3(tf / pk["tf_sustained"]),
            "projection_only_ms": r3(ms_proj), "library_linear_tanh_cosine_ms": r3(ms_lib),
            "train_step": {"workload": "train() mode, dropout 0.1: dropout + GEMM(bias, tanh, dropout) + fused cosine/bce loss writing d_pre + "
                                       "dgrad GEMM + MN-major wgrad GEMM + db, gradients for f1, f2, W, b",
                           "ms": r3(ms_train), "tf": r3(tf_train), "frac_sustained": r3(tf_train / pk["tf_sustained"]),
                           "library_autograd_ms": r3(ms_train_lib)}}


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
    ap.add_argument("--only-headline", "--no-retrieval", dest="only_headline", action="store_true",
                    help="config 2 only (skip configs 1, 3, 4, 5, softmax, projection)")
    ap.add_argument("--no-config5", action="store_true")
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
    # The K timed steps are captured into ONE CUDA graph and replayed: the eager call costs ~60 us of Python / allocator time
    # per 92 us step, so any host hiccup (this process also polls nvidia-smi for the clocks) shows up as GPU idle time between
    # launches.  Same launches, same buffers' worth of work per step.  IA_BENCH_GRAPH=0, or any capture failure: eager loop.
    graph, g, graph_launches = None, None, 0
    if os.environ.get("IA_BENCH_GRAPH", "1") != "0":
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                out = step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            lc = ia.launch_count()
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                for _ in range(args.steps):
                    out = step()
            graph_launches = ia.launch_count() - lc
            g.replay()                         # untimed: the graph's first launch uploads it
            torch.cuda.synchronize()
            graph = g
        except Exception as exc:               # noqa: BLE001 -- fall back to the eager loop, and say so
            sys.stderr.write(f"bench.py: CUDA-graph capture of the timed steps failed ({exc!r}); timing the eager loop\n")
            graph = None
            torch.cuda.synchronize()
            out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ia.launch_count()
    e0.record()
    if graph is not None:
        graph.replay()
    else:
        for _ in range(args.steps):
            out = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = graph_launches if graph is not None else ia.launch_count() - l0
    # the same K steps launched eagerly from Python, for comparison (reported, not the headline)
    for _ in range(3):                          # the caching allocator re-grows its ordinary pool after the capture
        out = step()
    torch.cuda.synchronize()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    for _ in range(args.steps):
        out = step()
    ee1.record()
    torch.cuda.synchronize()
    eager_ms_step = ee0.elapsed_time(ee1) / args.steps
    used_graph = graph is not None
    graph = g = None                            # releases the graph's private memory pool
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

    # ---------------------------------------------------------------- e2e: host buffers through the C ABI, gradients returned
    node = gpu_numa_node(torch, local_rank)
    with NumaPreferred(node) as numa:
        xh, yh, lh = x.cpu().pin_memory(), y.cpu().pin_memory(), labels.cpu().pin_memory()
        dxh, dyh = torch.empty_like(xh).pin_memory(), torch.empty_like(yh).pin_memory()
        loss_h = torch.zeros(1).pin_memory()

    def e2e_step():
        _lib.check(lib.ia_pair_score_loss_host(0, 0, 1.0, 1, _lib.IA_BF16, xh.data_ptr(), yh.data_ptr(), lh.data_ptr(), N_PAIRS, DIM,
                                               loss_h.data_ptr(), dxh.data_ptr(), dyh.data_ptr(), local_rank))
        return float(loss_h[0])
    for _ in range(3):
        e2e_loss = e2e_step()
    e2e_grads_equal = bool(torch.equal(dxh, out[3].cpu()) and torch.equal(dyh, out[4].cpu()))
    e2e_steps = max(3, min(args.steps, 50))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_loss = e2e_step()            # returns after loss, dx and dy are on the host (stream syncs inside)
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], device=device)
    my_e2e_ms = float(e2e_ms)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_val = N_PAIRS * world / (float(e2e_ms) * 1e-3)
    n_chunks = -(-N_PAIRS // max(1024, (16 << 20) // (DIM * 2)))
    h2d = 2 * N_PAIRS * DIM * 2 + N_PAIRS * 8
    d2h = 2 * N_PAIRS * DIM * 2 + 4 * n_chunks
    e2e = {"value": e2e_val, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": float(e2e_ms), "steps": e2e_steps,
           "api": "ia_pair_score_loss_host (pinned host x, y, labels -> loss, dx, dy in pinned host memory)", "loss": e2e_loss,
           "grads_equal_device_path": e2e_grads_equal,
           "h2d_GBs_this_rank": r3(h2d / my_e2e_ms / 1e6), "d2h_GBs_this_rank": r3(d2h / my_e2e_ms / 1e6),
           "host_memory": {"gpu_numa_node": node, "policy": numa.state}}
    del dxh, dyh

    # ---------------------------------------------------------------- the other BASELINE configs
    others, retrieval, c5, projection = None, None, None, None
    if not args.only_headline:
        del x, y, out
        torch.cuda.empty_cache()
        if rank == 0:
            others = bench_pair_configs(torch, F_, _lib, device, local_rank, pk)
        if world > 1:
            dist.barrier()
        retrieval = bench_retrieval(torch, dist, ia, device, rank, world, pk)
        torch.cuda.empty_cache()
        if not args.no_config5:
            c5 = bench_config5(torch, dist, ia, device, rank, world, pk)
        if rank == 0:
            torch.cuda.empty_cache()
            projection = bench_projection(torch, ia, device, pk)
        if world > 1:
            dist.barrier()
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
        summary = {"c2": {"Mpairs_s": r3(value / 1e6), "us": r3(ms_step * 1e3), "hbm_frac": r3(gbs / pk["hbm"]), "eager_us": r3(eager_ms_step * 1e3),
                          "e2e_Mpairs_s": r3(e2e_val / 1e6), "e2e_ms": r3(float(e2e_ms))}}
        if others:
            summary["c1"] = others["c1"]
            summary["c3"] = others["c3"]
            summary["softmax_ce"] = others["softmax_ce"]
        if retrieval:
            summary["c4"] = {"qps": r3(retrieval["value"]), "ms": r3(retrieval["ms_per_step"]), "n_gpus": world,
                             "tf_per_gpu": r3(retrieval["roofline"]["achieved"]), "frac_sustained": r3(retrieval["roofline"]["frac"]),
                             "parity_ok": retrieval["parity_ok"], "oracle_ok": bool(retrieval["oracle_slab"]["scores_ok"] and
                                                                                     retrieval["oracle_slab"]["rows_ok_where_separated"])}
        if c5:
            summary["c5"] = {"qps": r3(c5["qps"]), "ms": c5["ms_per_pass"], "rows_total": c5["rows_total"], "n_gpus": world,
                             "tf_per_gpu": c5["tf_per_gpu"], "frac_sustained": c5["frac_sustained"],
                             "oracle_ok": bool(c5["oracle_slab"]["scores_ok"] and c5["oracle_slab"]["rows_ok_where_separated"])}
        if projection:
            summary["projection"] = {"ms": projection["ms"], "frac_sustained": projection["frac_sustained"],
                                     "library_ms": projection["library_linear_tanh_cosine_ms"],
                                     "train_step_ms": projection["train_step"]["ms"], "train_step_library_ms": projection["train_step"]["library_autograd_ms"]}
        line = {
            "metric": "pairs/s (fused score+loss fwd/bwd)", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": config_of(world),
            "loss": loss_val, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "step_launch": "cuda_graph (the K timed steps captured once, one replay timed)" if used_graph else "eager",
            "eager_ms_per_step": eager_ms_step,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "traffic": traffic,
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": pk["source"] + " (copy bandwidth)",
                         "frac_of_nominal_8000": gbs / 8000.0, "per_gpu": True},
            "cpu_baseline": cpu, "projection": projection, "retrieval": retrieval, "config5": c5,
            "summary": summary,          # LAST: every BASELINE config, compact, inside the tail the driver keeps
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
