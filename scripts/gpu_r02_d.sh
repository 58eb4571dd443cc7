#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:softmax_head_mma_kernel -s 2 -c 1 -o gpurun_out/ncu_softmax_mma -f python scripts/prof_softmax.py bf16 > gpurun_out/ncu_softmax_mma.log 2>&1; tail -n 1 gpurun_out/ncu_softmax_mma.log
