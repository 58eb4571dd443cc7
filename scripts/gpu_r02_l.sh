#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pair.py -m gpu -q -k softmax 2>&1 | tail -n 3
echo "== 32 pairs per stage at h = 512 (default)"; timeout 300 python scripts/bench_softmax.py 2>&1 | tee gpurun_out/softmax_rb32.log
echo "== 16 pairs per stage everywhere"; IA_HEAD_RB16=1 timeout 300 python scripts/bench_softmax.py 2>&1 | grep "h=  512" | tee gpurun_out/softmax_rb16.log
