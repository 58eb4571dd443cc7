#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q --timeout 600 -k "pair or probed or shard_merge" 2>&1 | tail -n 2
timeout 300 python scripts/exp_retrieval_pair.py > $O/retrieval_pair_after_scope_fix.log 2>&1; tail -n 16 $O/retrieval_pair_after_scope_fix.log
WAIT_FLAGS=142,206 timeout 200 python scripts/exp_retrieval_mma_waits.py 10 2>&1 | tail -n 4 | tee $O/retrieval_pair_waits_after_scope_fix.log
WAIT_FLAGS=142 timeout 200 python scripts/exp_retrieval_mma_waits.py 100 2>&1 | tail -n 2 | tee -a $O/retrieval_pair_waits_after_scope_fix.log
IA_RETR_PAIR=1 timeout 300 ncu --set full --clock-control none -k regex:retrieve_tc_kernel -s 1 -c 1 -o $O/ncu_retr_pair1_fixed -f python scripts/prof_retrieval.py 4096 500000 1024 cosine 10 2 > $O/ncu_retr_pair1_fixed.log 2>&1; tail -n 1 $O/ncu_retr_pair1_fixed.log
