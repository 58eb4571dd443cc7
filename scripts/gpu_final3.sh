#!/bin/bash
# end-of-round evidence after the last kernel changes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -n 1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 1
timeout 600 python bench.py > gpurun_out/bench_n1_r01d.json 2> gpurun_out/bench_n1_r01d.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_r01d.json 2>/dev/null; echo "reference rc $?"
timeout 200 python scripts/bench_pair_configs.py > gpurun_out/pair_configs_r01d.log 2>&1; tail -n 9 gpurun_out/pair_configs_r01d.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:retrieve_simt_kernel -c 1 -o gpurun_out/ncu_simt_l2 -f python scripts/prof_retrieval.py 1024 65536 1024 l2 100 2 > gpurun_out/ncu_simt_l2.log 2>&1; tail -n 1 gpurun_out/ncu_simt_l2.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:softmax_head_kernel -s 4 -c 1 -o gpurun_out/ncu_softmax_12w -f python scripts/prof_softmax.py bf16 > gpurun_out/ncu_softmax_12w.log 2>&1; tail -n 1 gpurun_out/ncu_softmax_12w.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launch_list_bench_r01d.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; grep -c "ia::" gpurun_out/ncu_launch_list_bench_r01d.csv
