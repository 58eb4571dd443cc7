#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -2
echo "== pair tests with the bulk ring on"; IA_PAIR_BULK=1 timeout 900 python -m pytest tests/test_gpu_pair.py -m gpu -q --timeout 600 2>&1 | tail -1
echo "== l1 / l2 / fp32 retrieval on the CUDA-core kernel"
for args in "2048 262144 1024 l2 100 4" "2048 262144 1024 l1 100 4" "2048 262144 512 l2 10 4"; do timeout 300 python scripts/prof_retrieval.py $args 2>&1 | tail -n 1; done | tee gpurun_out/retr_simt.log
echo "== sanitizer"; timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py 2>&1 | tail -1
