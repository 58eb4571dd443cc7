#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_pair.py tests/test_gpu_projection.py tests/test_gpu_next_rows.py -m gpu -q --timeout 600 -x 2>&1 | tail -n 3
for pdl in 0 1 0 1; do echo "== IA_PDL=$pdl"; IA_PDL=$pdl timeout 200 python scripts/bench_pair_configs.py 2>&1 | grep -E "^C[123]|softmax"; done | tee $O/pair_configs_pdl_ab.log
for pdl in 0 1; do echo "== IA_PDL=$pdl"; IA_PDL=$pdl timeout 200 python scripts/bench_softmax.py 2>&1 | cut -c1-150; done | tee $O/softmax_pdl_ab.log
