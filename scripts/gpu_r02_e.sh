#!/bin/bash
# round 2, session E: full parity suite + smoke + bench (N=1) + racecheck split by kernel family
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -n 4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 1
timeout 900 python bench.py > gpurun_out/bench_n1_r02e.json 2> gpurun_out/bench_n1_r02e.err; echo "bench rc $?"; tail -c 1300 gpurun_out/bench_n1_r02e.json
# racecheck: (a) every kernel that synchronises with bar.sync / shuffles only (must be clean), (b) the mbarrier / TMA / tcgen05 kernels
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 2000000 --kernel-name-exclude regex:"retrieve_tc_kernel|project_kernel|softmax_head_mma_kernel" python scripts/sanitize_small.py > /tmp/race_a.log 2>&1
python scripts/racecheck_summary.py /tmp/race_a.log > gpurun_out/racecheck_r02_barsync_kernels.txt; tail -n 3 gpurun_out/racecheck_r02_barsync_kernels.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 2000000 --kernel-name regex:"retrieve_tc_kernel|project_kernel|softmax_head_mma_kernel" python scripts/sanitize_small.py > /tmp/race_b.log 2>&1
python scripts/racecheck_summary.py /tmp/race_b.log > gpurun_out/racecheck_r02_mbarrier_kernels.txt; head -n 40 gpurun_out/racecheck_r02_mbarrier_kernels.txt | cut -c1-200
