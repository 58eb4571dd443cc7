#!/bin/bash
# packed-fp32 CUDA-core retrieval, warp-role order A/B (IA_RETR_FLAGS bit 4), projection re-bench
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_catalog_file.py tests/test_gpu_projection.py -m gpu -q -x --timeout 600 2>&1 | tail -n 3
echo "== l1/l2 CUDA-core kernel"
for args in "2048 262144 1024 l2 100" "2048 262144 1024 l1 100"; do timeout 300 python scripts/prof_retrieval.py $args 2>&1 | tail -n 1; done
echo "== role order A/B"
for f in 14 30 14 30; do
  echo "IA_RETR_FLAGS=$f"
  IA_RETR_FLAGS=$f timeout 300 python scripts/prof_retrieval.py 10000 1000000 1024 cosine 100 2>&1 | tail -n 1
  IA_RETR_FLAGS=$f timeout 300 python scripts/prof_retrieval.py 10000 1000000 1024 cosine 10 2>&1 | tail -n 1
done
echo "== projection"
timeout 200 python scripts/bench_projection.py 2>&1 | tail -n 1
