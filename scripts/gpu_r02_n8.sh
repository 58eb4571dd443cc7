#!/bin/bash
# round 2: 8-GPU session -- exchange-strategy A/B for C4, then bench at N = 8 (the driver's command line) and N = 1 on the same box
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 scripts/bench_c4_sharded.py 2>&1 | grep "^N=" | tee gpurun_out/c4_sharded_ab_n8_r02.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 scripts/bench_c4_sharded.py 2>&1 | grep "^N=" | tee gpurun_out/c4_sharded_ab_n4_r02.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_r02.json 2> gpurun_out/bench_n8_r02.err; echo "bench N=8 rc $?"
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
