#!/bin/bash
# checkpoint run: all GPU tests, smoke, bench (both arms), launch-gap experiment
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/t_gpu.log 2>&1; echo "exit $?"; tail -n 6 gpurun_out/t_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -n 2 gpurun_out/smoke.log
echo "== gap experiment"; timeout 300 python scripts/exp_pair_gap.py 2>&1 | tail -n 3 | tee gpurun_out/pair_gap.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "exit $?"; cat gpurun_out/bench_ref.json
