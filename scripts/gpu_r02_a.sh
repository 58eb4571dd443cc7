#!/bin/bash
# round 2, session A: parity suite (incl. the new full-size oracle comparisons), smoke, bench N=1 + reference arm, racecheck
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -n 15
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2
timeout 900 python bench.py > gpurun_out/bench_n1_r02a.json 2> gpurun_out/bench_n1_r02a.err; echo "bench rc $?"; tail -c 1600 gpurun_out/bench_n1_r02a.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_r02a.json 2>/dev/null; echo "reference rc $?"
timeout 200 python scripts/bench_softmax.py > gpurun_out/softmax_r02a.log 2>&1; cat gpurun_out/softmax_r02a.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_small.py > gpurun_out/racecheck_r02a.log 2>&1; tail -n 5 gpurun_out/racecheck_r02a.log
