#!/bin/bash
# round checkpoint: all GPU tests, smoke, bench (both arms), ncu evidence for the two dominant kernels
mkdir -p gpurun_out
echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/t_gpu.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/t_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/smoke.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cut -c1-400 gpurun_out/bench.json; tail -n 2 gpurun_out/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "exit $?"
echo "== pair configs"; python scripts/bench_pair_configs.py | tee gpurun_out/pair_configs.log
echo "== ncu launch list of the default bench command (short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "exit $?"
echo "== ncu full: fused pair kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 3 -c 2 -o gpurun_out/prof_pair -f python bench.py --steps 5 --warmup 3 --no-retrieval --no-cpu-baseline > gpurun_out/ncu_pair.log 2>&1; echo "exit $?"
echo "== ncu full: retrieval tc kernel (C4 shape)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:retrieve_tc -s 1 -c 1 -o gpurun_out/prof_retr10k -f python scripts/prof_retrieval.py 10000 1000000 1024 cosine 100 3 > gpurun_out/ncu_retr.log 2>&1; echo "exit $?"
