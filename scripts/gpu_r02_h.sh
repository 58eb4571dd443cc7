#!/bin/bash
# round 2, session H (one GPU): the evidence set -- parity suite, smoke, bench + reference arm, per-kernel timing logs, ncu captures
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -n 2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 1
timeout 900 python bench.py > $O/bench_n1_r02h.json 2> $O/bench_n1_r02h.err; echo "bench rc $?"; tail -c 1500 $O/bench_n1_r02h.json; echo
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference_r02h.json 2>/dev/null; echo "reference rc $?"
timeout 200 python scripts/bench_pair_configs.py > $O/pair_configs_r02h.log 2>&1; tail -n 10 $O/pair_configs_r02h.log
timeout 200 python scripts/bench_softmax.py > $O/softmax_r02h.log 2>&1; cat $O/softmax_r02h.log
IA_HEAD_MMA=0 timeout 200 python scripts/bench_softmax.py > $O/softmax_cudacore_r02h.log 2>&1; cat $O/softmax_cudacore_r02h.log
timeout 300 python scripts/exp_retrieval_probe.py > $O/retrieval_probe_r02h.log 2>&1; cat $O/retrieval_probe_r02h.log
timeout 200 python scripts/prof_shard_step.py 8 > $O/shard_step_r02h.log 2>&1; cat $O/shard_step_r02h.log
timeout 200 python scripts/prof_shard_step.py 8 0.03125 2>&1 | head -n 6 | tee -a $O/shard_step_r02h.log
timeout 200 python scripts/prof_shard_step.py 8 0.015625 2>&1 | head -n 6 | tee -a $O/shard_step_r02h.log
timeout 200 python scripts/prof_shard_step.py 4 0.03125 2>&1 | head -n 6 | tee -a $O/shard_step_r02h.log
timeout 200 python scripts/prof_shard_step.py 2 0.03125 2>&1 | head -n 6 | tee -a $O/shard_step_r02h.log
timeout 200 python scripts/prof_projection_train.py 6 2>/dev/null
timeout 200 python scripts/bench_projection.py > $O/projection_bench_r02h.log 2>&1; tail -n 8 $O/projection_bench_r02h.log
# launch lists (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/ncu_launch_list_bench_r02h.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-config5 > /dev/null 2>&1; grep -c "ia::" $O/ncu_launch_list_bench_r02h.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ncu_launch_list_projection_train_r02h.csv python scripts/prof_projection_train.py 3 > /dev/null 2>&1
# ncu --set full captures of the round-2 kernels
timeout 400 ncu --set full --clock-control none --import-source on -k regex:softmax_head_mma_kernel -s 2 -c 1 -o $O/ncu_softmax_mma_r02h -f python scripts/prof_softmax.py bf16 > $O/ncu_softmax_mma_r02h.log 2>&1; tail -n 1 $O/ncu_softmax_mma_r02h.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 1 -c 1 -o $O/ncu_wgrad_r02h -f python scripts/prof_projection_train.py 3 > $O/ncu_wgrad_r02h.log 2>&1; tail -n 1 $O/ncu_wgrad_r02h.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:project_kernel -s 2 -c 2 -o $O/ncu_project_train_r02h -f python scripts/prof_projection_train.py 3 > $O/ncu_project_train_r02h.log 2>&1; tail -n 1 $O/ncu_project_train_r02h.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:retrieve_tc_kernel -s 4 -c 2 -o $O/ncu_retrieve_tc_r02h -f python scripts/prof_retrieval.py 10000 1000000 1024 cosine 100 3 > $O/ncu_retrieve_tc_r02h.log 2>&1; tail -n 1 $O/ncu_retrieve_tc_r02h.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 5 -c 1 -o $O/ncu_pair_r02h -f python scripts/exp_pair_gap.py > $O/ncu_pair_r02h.log 2>&1; tail -n 1 $O/ncu_pair_r02h.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:radix_scatter_kernel -s 1 -c 1 -o $O/ncu_radix_scatter_r02h -f python -c "
import sys; sys.path.insert(0, '.')
import torch, item_alignment_b200 as ia
s = torch.rand(10_000_000, device='cuda'); l = (torch.rand(10_000_000, device='cuda') < 0.3).long()
print(ia.find_best_f1_and_threshold(s, l))" > $O/ncu_radix_scatter_r02h.log 2>&1; tail -n 1 $O/ncu_radix_scatter_r02h.log
ls -la $O/*.ncu-rep | tail -n 8
