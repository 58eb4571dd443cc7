#!/bin/bash
# round 2, session B: tensor-core softmax head (parity + timing A/B), host pipeline chunk A/B, racecheck by kernel family
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_next_rows.py -m gpu -x -q --timeout 600 -k "softmax or gather or ladder or upstream" 2>&1 | tail -n 12
echo "--- softmax head, tensor-core kernel"; timeout 200 python scripts/bench_softmax.py 2>&1 | tail -n 6
echo "--- softmax head, CUDA-core kernel (IA_HEAD_MMA=0)"; IA_HEAD_MMA=0 timeout 200 python scripts/bench_softmax.py 2>&1 | tail -n 6
for mb in 2 4 8 16 32; do IA_HOST_CHUNK_MB=$mb timeout 120 python scripts/exp_host_chunk.py 2>&1 | tail -n 1; done
