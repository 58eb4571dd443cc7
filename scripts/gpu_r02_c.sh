#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py -m gpu -x -q --timeout 300 -k "softmax or upstream" 2>&1 | tail -n 3
for dbg in 0 4 8 12; do echo "--- IA_HEAD_DEBUG=$dbg"; IA_HEAD_DEBUG=$dbg timeout 200 python scripts/bench_softmax.py 2>&1 | grep -v float32; done
