#!/bin/bash
mkdir -p gpurun_out
echo "== pair tests (bulk default)"; timeout 1200 python -m pytest tests/test_gpu_pair.py -m gpu -q -x --timeout 600 > gpurun_out/t_pair.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/t_pair.log
echo "== pair tests (bulk off)"; IA_PAIR_BULK=0 timeout 1200 python -m pytest tests/test_gpu_pair.py -m gpu -q -x --timeout 600 > gpurun_out/t_pair0.log 2>&1; echo "exit $?"; tail -n 2 gpurun_out/t_pair0.log
for b in 1 0 1 0; do IA_PAIR_BULK=$b python scripts/bench_pair_configs.py; done | tee gpurun_out/pair_configs.log
