#!/bin/bash
mkdir -p gpurun_out
echo "== retrieval tests"; timeout 1500 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x --timeout 900 > gpurun_out/t_retr.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/t_retr.log
for flags in 14 6; do
for args in "10000 1000000 1024 cosine 100" "10000 1000000 1024 cosine 10" "2048 1000000 1024 cosine 100"; do
    IA_RETR_FLAGS=$flags timeout 300 python scripts/prof_retrieval.py $args 2>&1 | tail -n 2; echo "  [flags $flags]"
done
done | tee gpurun_out/retr_ab.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:retrieve_tc -s 1 -c 1 -o gpurun_out/prof_retr10k -f python scripts/prof_retrieval.py 10000 1000000 1024 cosine 100 3 > gpurun_out/ncu_retr.log 2>&1; echo "ncu exit $?"
