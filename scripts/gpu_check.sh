#!/bin/bash
# One gpurun call: parity tests, smoke, bench, ncu launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
echo "== pair tests"; timeout 1200 python -m pytest tests/test_gpu_pair.py -m gpu -q -x --timeout 600 > gpurun_out/t_pair.log 2>&1; echo "exit $?"; tail -n 30 gpurun_out/t_pair.log
echo "== retrieval exact"; timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "exact_fixture or k_larger or merge_and_unpack" --timeout 300 > gpurun_out/t_retr1.log 2>&1; echo "exit $?"; tail -n 40 gpurun_out/t_retr1.log
echo "== retrieval rest"; timeout 1500 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "not (exact_fixture or k_larger or merge_and_unpack)" --timeout 900 > gpurun_out/t_retr2.log 2>&1; echo "exit $?"; tail -n 40 gpurun_out/t_retr2.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/smoke.log
echo "== retrieval perf"; timeout 300 python scripts/prof_retrieval.py 2048 1000000 1024 cosine > gpurun_out/retr_perf.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/retr_perf.log
echo "== bench"; timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit $?"; tail -n 3 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
