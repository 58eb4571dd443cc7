#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_projection.py tests/test_gpu_pair.py -m gpu -q --timeout 300 -k "train or project or vecsim or ladder" 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-300
timeout 600 python bench.py --no-config5 > gpurun_out/bench_n1_r02f.json 2> gpurun_out/bench_n1_r02f.err; echo "bench rc $?"; tail -c 900 gpurun_out/bench_n1_r02f.json; tail -n 5 gpurun_out/bench_n1_r02f.err
IN="--kernel-name kns=retrieve_tc_kernel --kernel-name kns=project_kernel --kernel-name kns=softmax_head_mma_kernel --kernel-name kns=wgrad_kernel"
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 3 $IN python scripts/sanitize_small.py > /tmp/race_b.log 2>&1; grep -v "^=========     \|Host Frame" /tmp/race_b.log | tail -n 25 | cut -c1-250
