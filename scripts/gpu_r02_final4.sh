#!/bin/bash
# round 2, last single-GPU records: bench (timed steps replayed from a CUDA graph) + reference arm + launch list of the same command
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py > $O/bench_n1_final4_r02.json 2> $O/bench_n1_final4_r02.err; echo "bench rc $?"; tail -c 1600 $O/bench_n1_final4_r02.json; echo; tail -n 3 $O/bench_n1_final4_r02.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference_final4_r02.json 2>/dev/null; echo "reference rc $?"
timeout 300 python bench.py --steps 20 --warmup 5 --only-headline --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('driver-style run (steps 20, warmup 5):', round(d['ms_per_step']*1e3,2), 'us', round(d['roofline']['frac'],4), d['step_launch'], 'eager', round(d['eager_ms_per_step']*1e3,2))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/ncu_launch_list_bench_final4_r02.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-config5 > /dev/null 2>&1; grep -c "ia::" $O/ncu_launch_list_bench_final4_r02.csv
timeout 300 python -m pytest tests -m gpu -q --timeout 600 -k "abi or smoke or host or graft" 2>&1 | tail -n 2
