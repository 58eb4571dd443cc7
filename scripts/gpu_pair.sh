#!/bin/bash
mkdir -p gpurun_out
echo "== pair tests"; timeout 1200 python -m pytest tests/test_gpu_pair.py -m gpu -q -x --timeout 600 > gpurun_out/t_pair.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/t_pair.log
for r in 1 0 1 0; do echo "IA_PAIR_ROWS2=$r"; IA_PAIR_ROWS2=$r python scripts/bench_pair_configs.py; done | tee gpurun_out/pair_configs.log
