#!/bin/bash
# round 2: 2-GPU session -- NCCL sharded retrieval test, bench at N = 2 (sharded parity_ok, config 5 with 2 shards, e2e per rank)
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -n 12
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q --timeout 300 -k "two_gpu" 2>&1 | tail -n 3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_r02.json 2> gpurun_out/bench_n2_r02.err; echo "bench N=2 rc $?"
tail -c 1500 gpurun_out/bench_n2_r02.json; echo; grep -i "error\|Traceback" -A5 gpurun_out/bench_n2_r02.err | head -n 30
python -c "
import json
d=json.load(open('gpurun_out/bench_n2_r02.json'))
print(json.dumps(d['e2e'])); print(json.dumps(d['retrieval'])[:1200])"
timeout 300 python scripts/prof_projection_train.py 4
