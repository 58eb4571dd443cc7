#!/bin/bash
# round 2, closing single-GPU evidence after programmatic dependent launch, the pair retrieval form and the max-filter epilogue
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -n 2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 1
timeout 900 python bench.py > $O/bench_n1_final3_r02.json 2> $O/bench_n1_final3_r02.err; echo "bench rc $?"; tail -c 1500 $O/bench_n1_final3_r02.json; echo
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference_final3_r02.json 2>/dev/null; echo "reference rc $?"
timeout 200 python scripts/bench_pair_configs.py > $O/pair_configs_final3_r02.log 2>&1; tail -n 10 $O/pair_configs_final3_r02.log
timeout 200 python scripts/bench_softmax.py > $O/softmax_head_mma_final3.log 2>&1; cut -c1-150 $O/softmax_head_mma_final3.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 5 -c 1 -o $O/ncu_pair_final3_r02 -f python scripts/exp_pair_gap.py > $O/ncu_pair_final3_r02.log 2>&1; tail -n 1 $O/ncu_pair_final3_r02.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/ncu_launch_list_bench_final3_r02.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-config5 > /dev/null 2>&1; grep -c "ia::" $O/ncu_launch_list_bench_final3_r02.csv
