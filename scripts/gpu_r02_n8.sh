#!/bin/bash
# round 2: 8-GPU session -- bench at N = 8 (the driver's command line), N = 1 on the same box for the scaling ratio
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -n 12 | cut -c1-160
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_r02.json 2> gpurun_out/bench_n8_r02.err; echo "bench N=8 rc $?"
tail -c 1500 gpurun_out/bench_n8_r02.json; echo; grep -i "error\|Traceback" -A5 gpurun_out/bench_n8_r02.err | head -n 30
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_8gpubox_r02.json 2> gpurun_out/bench_n1_8gpubox_r02.err; echo "bench N=1 rc $?"
python - <<'PY'
import json
a=json.load(open('gpurun_out/bench_n8_r02.json')); b=json.load(open('gpurun_out/bench_n1_8gpubox_r02.json'))
print('C4 N=8', a['retrieval']['ms_per_step'], 'ms; N=1', b['retrieval']['ms_per_step'], 'ms; ratio', b['retrieval']['ms_per_step']/a['retrieval']['ms_per_step'])
print('parity', a['retrieval']['parity_ok'], a['retrieval']['oracle_slab'])
print('C5 N=8', a['config5'])
print('e2e N=8', a['e2e']['value'], a['e2e']['ms_per_step'], a['e2e']['h2d_GBs_this_rank'], 'N=1', b['e2e']['value'])
print('pair N=8', a['value'], a['ms_per_step'], 'N=1', b['value'], b['ms_per_step'])
PY
