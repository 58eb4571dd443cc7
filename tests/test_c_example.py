"""The C ABI from a plain C program (examples/pair_score_host.c): no Python, no torch in the calling process.
CPU: the example compiles against include/ia_b200.h with gcc (the header is valid C) and links to the library.
GPU: it runs, and its output equals an independent numpy recomputation of the same generated inputs."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "item_alignment_b200")


def build_example(tmp_path):
    exe = str(tmp_path / "pair_score_host")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "pair_score_host.c"),
           "-o", exe, "-L", LIBDIR, "-lia_b200", f"-Wl,-rpath,{LIBDIR}", "-lm"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    return exe


def test_c_example_compiles_and_links(tmp_path):
    if not os.path.isfile(os.path.join(LIBDIR, "libia_b200.so")):
        pytest.skip("library not built")
    exe = build_example(tmp_path)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libia_b200.so" in out and "not found" not in out.split("libia_b200.so")[1].splitlines()[0]
    assert "libtorch" not in out and "libpython" not in out


def lcg_inputs(n, d):
    state = 20221009
    def nxt():
        nonlocal state
        state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
        return np.float32(state >> 8) * np.float32(1.0 / 8388608.0) - np.float32(1.0)
    x = np.array([nxt() for _ in range(n * d)], dtype=np.float32)
    y = np.empty(n * d, dtype=np.float32)
    for i in range(n * d):
        y[i] = x[i] if i % 3 == 0 else nxt()
    return x.reshape(n, d), y.reshape(n, d)


@pytest.mark.gpu
@pytest.mark.parametrize("measure", ["cosine", "l2", "inner_product"])
def test_c_example_runs_and_matches_numpy(tmp_path, measure):
    from oracle import formula
    exe = build_example(tmp_path)
    n, d = 600, 48
    proc = subprocess.run([exe, measure, str(n), str(d)], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr
    head = proc.stdout.splitlines()[0]
    m = re.search(r"positives=(\d+) sum_sim=(\S+) sum_probs=(\S+) launches=(\d+)", head)
    assert m and int(m.group(4)) >= 1, head
    x, y = lcg_inputs(n, d)
    s = formula.score(measure, x, y)
    p = formula.probs(measure, s)
    scale = float(np.abs(s).sum()) + 1.0
    assert abs(float(m.group(2)) - float(s.sum())) <= 1e-5 * scale
    assert abs(float(m.group(3)) - float(p.sum())) <= 1e-5 * n
    near = int((np.abs(p - 0.5) < 1e-6).sum())                         # pairs sitting on the threshold may go either way
    assert abs(int(m.group(1)) - int((p >= 0.5).sum())) <= near
    for line in proc.stdout.splitlines()[1:]:
        i, sim = int(line.split()[1]), float(re.search(r"sim=(\S+)", line).group(1))
        assert abs(sim - float(s[i])) <= 1e-5 * (abs(float(s[i])) + float(np.abs(x[i]).dot(np.abs(y[i]))))
    bad = subprocess.run([exe, "softmax", "10", "8"], capture_output=True, text=True)
    assert bad.returncode == 2 and "Unsupported similarty measure" in bad.stderr
