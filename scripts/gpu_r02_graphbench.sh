#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for k in "20 5" "200 10"; do set -- $k; timeout 300 python bench.py --only-headline --no-cpu-baseline --steps $1 --warmup $2 2> $O/bench_graph_$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('steps', d['steps'], 'ms_per_step', round(d['ms_per_step']*1e3,2), 'us  frac', round(d['roofline']['frac'],4), ' eager', round(d['eager_ms_per_step']*1e3,2), 'us  launch:', d['step_launch'], ' gpu_launches', d['gpu_launches'], ' loss', d['loss'])"; cat $O/bench_graph_$1.err | tail -n 2; done
IA_BENCH_GRAPH=0 timeout 300 python bench.py --only-headline --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('IA_BENCH_GRAPH=0: steps', d['steps'], 'ms_per_step', round(d['ms_per_step']*1e3,2), 'us  launch:', d['step_launch'])"
