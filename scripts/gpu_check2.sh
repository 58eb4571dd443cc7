#!/bin/bash
# retrieval iteration: parity tests + perf sweep + ncu captures
mkdir -p gpurun_out
echo "== retrieval tests"; timeout 1500 python -m pytest tests/test_gpu_retrieval.py -m gpu -q --timeout 900 > gpurun_out/t_retr.log 2>&1; echo "exit $?"; tail -n 15 gpurun_out/t_retr.log
echo "== retrieval perf"
for args in "2048 1000000 1024 cosine 100" "2048 1000000 1024 cosine 1" "2048 1000000 1024 inner_product 100" "10000 1000000 1024 cosine 100" "10000 1000000 1024 cosine 10" "8192 1000000 512 inner_product 10"; do
  timeout 300 python scripts/prof_retrieval.py $args 2>&1 | tail -n 1
done | tee gpurun_out/retr_perf.log
if [ "$1" = "ncu" ]; then
echo "== ncu launch list (bench, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "exit $?"
echo "== ncu full: fused pair kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 3 -c 2 -o gpurun_out/prof_pair -f python bench.py --steps 5 --warmup 3 --no-retrieval --no-cpu-baseline > gpurun_out/ncu_pair.log 2>&1; echo "exit $?"
echo "== ncu full: retrieval tc kernel"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:retrieve_tc -c 1 -o gpurun_out/prof_retr -f python scripts/prof_retrieval.py 2048 262144 1024 cosine 100 > gpurun_out/ncu_retr.log 2>&1; echo "exit $?"
ls -la gpurun_out
fi
