"""CPU: the reference arm of bench.py honours the driver contract (exactly one JSON line on stdout, required keys)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]


def test_bench_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                          timeout=300, cwd=ROOT)
    assert proc.returncode != 0 and "no CPU path" in (proc.stderr + proc.stdout)


import pytest  # noqa: E402


@pytest.mark.gpu
def test_native_arm_headline_line_on_gpu():
    """GPU: the headline leg of bench.py (config 2 only) prints one JSON line with the contract's keys; the K timed steps are
    replayed from a CUDA graph, one launch of the library's own kernel per step, gradients of the host-buffer path equal to
    those of the device path."""
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--only-headline", "--no-cpu-baseline", "--steps", "6",
                           "--warmup", "3"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "clocks", "e2e", "gpu_launches", "step_launch", "eager_ms_per_step"):
        assert key in d, key
    assert d["steps"] == 6 and d["gpu_launches"] == 6 and d["n_gpus"] == 1 and d["dtype"] == "bf16"
    assert d["step_launch"].startswith("cuda_graph") or d["step_launch"] == "eager", d["step_launch"]   # eager: capture fell back
    assert 0.5 < d["roofline"]["frac"] < 1.05 and d["roofline"]["bound"] == "hbm"
    assert d["e2e"]["grads_equal_device_path"] is True
    assert d["e2e"]["h2d_bytes_per_step"] > 2.6e8 and d["e2e"]["d2h_bytes_per_step"] > 2.6e8
    assert abs(d["value"] - 65536 / (d["ms_per_step"] * 1e-3)) < 1e-3 * d["value"]
