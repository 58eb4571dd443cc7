#!/bin/bash
# end-of-round evidence: full GPU tests, smoke, bench line, launch list, ncu of the projection kernel, pair configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -n 1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 1
timeout 600 python bench.py > gpurun_out/bench_n1_r01c.json 2> gpurun_out/bench_n1_r01c.err; echo "bench rc $?"
timeout 200 python scripts/bench_pair_configs.py > gpurun_out/pair_configs_r01c.log 2>&1; tail -n 9 gpurun_out/pair_configs_r01c.log
timeout 200 python scripts/bench_projection.py > gpurun_out/projection_bench_r01c.log 2>&1; tail -n 1 gpurun_out/projection_bench_r01c.log | cut -c1-700
timeout 400 ncu --set full --clock-control none --import-source on -k regex:project_kernel -s 2 -c 2 -o gpurun_out/ncu_projection_final -f python scripts/prof_projection.py > gpurun_out/ncu_projection_final.log 2>&1; tail -n 1 gpurun_out/ncu_projection_final.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launch_list_bench_r01c.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; grep -c "ia::" gpurun_out/ncu_launch_list_bench_r01c.csv
