#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py -m gpu -x -q --timeout 300 -k "softmax or upstream or ladder" 2>&1 | tail -n 2
IA_HEAD_DEBUG=16 timeout 200 python scripts/exp_softmax_stats.py 2>&1 | tail -n 4
timeout 200 python scripts/bench_softmax.py 2>&1 | grep -v float32 | cut -c1-260
