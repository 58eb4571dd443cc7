#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q --timeout 300 -k "two_gpu" 2>&1 | tail -n 2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-config5 > gpurun_out/bench_n2_r02j.json 2> gpurun_out/bench_n2_r02j.err; echo "bench N=2 rc $?"
python -c "
import json
d=json.load(open('gpurun_out/bench_n2_r02j.json'))
print(d['summary']['c4'])"
