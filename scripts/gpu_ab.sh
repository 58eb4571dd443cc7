#!/bin/bash
mkdir -p gpurun_out
echo "== retrieval tests"; timeout 1500 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x --timeout 900 > gpurun_out/t_retr.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/t_retr.log
for args in "10000 1000000 1024 cosine 100" "10000 1000000 1024 cosine 10" "2048 1000000 1024 cosine 100" "10000 125000 1024 cosine 100" "10000 1000000 1024 inner_product 100" "8192 1000000 512 inner_product 10"; do
    timeout 300 python scripts/prof_retrieval.py $args 2>&1 | tail -n 2
done | tee gpurun_out/retr_ab.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_retr.csv python scripts/prof_retrieval.py 10000 125000 1024 cosine 100 3 > /dev/null 2>&1; grep -E "retrieve|merge|inv_norm" gpurun_out/launches_retr.csv | cut -d, -f5,15- | tail -6
