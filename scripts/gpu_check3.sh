#!/bin/bash
mkdir -p gpurun_out
echo "== retrieval tests"; timeout 1500 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x --timeout 900 > gpurun_out/t_retr.log 2>&1; echo "exit $?"; tail -n 15 gpurun_out/t_retr.log
echo "== retrieval perf"
for args in "2048 1000000 1024 cosine 100" "10000 1000000 1024 cosine 100" "10000 1000000 1024 cosine 10" "10000 1000000 1024 inner_product 100" "8192 1000000 512 inner_product 10" "2048 262144 1024 cosine 100"; do
  timeout 300 python scripts/prof_retrieval.py $args 2>&1 | tail -n 2
done | tee gpurun_out/retr_perf.log
if [ "$1" = "ncu" ]; then
echo "== ncu full: retrieval tc kernel 10k"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:retrieve_tc -s 1 -c 1 -o gpurun_out/prof_retr10k -f python scripts/prof_retrieval.py 10000 1000000 1024 cosine 100 > gpurun_out/ncu_retr.log 2>&1; echo "exit $?"
fi
