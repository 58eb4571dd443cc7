#!/bin/bash
# round 2, closing 8-GPU check: the driver's bench command at N = 8 and a short N = 1 run on the same box
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_final_r02.json 2> gpurun_out/bench_n8_final_r02.err; echo "bench N=8 rc $?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-config5 > gpurun_out/bench_n1_8gpubox_final_r02.json 2> gpurun_out/bench_n1_8gpubox_final_r02.err; echo "bench N=1 rc $?"
python - <<'PY'
import json
a=json.loads(open('gpurun_out/bench_n8_final_r02.json').read().strip().splitlines()[-1]); b=json.loads(open('gpurun_out/bench_n1_8gpubox_final_r02.json').read().strip().splitlines()[-1])
print('C4 N=8', a['retrieval']['ms_per_step'], 'ms; N=1', b['retrieval']['ms_per_step'], 'ms; ratio', b['retrieval']['ms_per_step']/a['retrieval']['ms_per_step'])
print('parity', a['retrieval']['parity_ok'], a['retrieval']['oracle_slab'])
print('C5 N=8', a['config5'])
print('e2e N=8', a['e2e']['value'], a['e2e']['ms_per_step'], 'N=1', b['e2e']['value'])
print('pair N=8', a['value'], a['ms_per_step'], a['step_launch'], 'eager', a['eager_ms_per_step'], '| N=1', b['value'], b['ms_per_step'])
PY
grep -i "capture\|Traceback" -A3 gpurun_out/bench_n8_final_r02.err | head -n 10
