#!/bin/bash
timeout 300 python scripts/exp_e2e_wc.py 2>&1 | grep "^N="
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 scripts/exp_e2e_wc.py 2>&1 | grep "^N="
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29556 scripts/exp_e2e_wc.py 2>&1 | grep "^N="
