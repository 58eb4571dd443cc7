#!/bin/bash
mkdir -p gpurun_out
echo "== retrieval tests"; timeout 1500 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x --timeout 900 > gpurun_out/t_retr.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/t_retr.log
for args in "10000 1000000 1024 cosine 100" "10000 1000000 1024 cosine 10" "2048 1000000 1024 cosine 100" "8192 1000000 512 inner_product 10" "10000 7812 1024 cosine 13"; do
    timeout 300 python scripts/prof_retrieval.py $args 2>&1 | tail -n 2
done | tee gpurun_out/retr_ab.log
python scripts/prof_shard_step.py 8 0.0625 | tee gpurun_out/shard_step.log
python scripts/prof_shard_step.py 4 0.0625 | tee -a gpurun_out/shard_step.log
python scripts/prof_shard_step.py 2 0.0625 | tee -a gpurun_out/shard_step.log
