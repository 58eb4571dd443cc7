#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for pair in 0 1; do for flags in 14 78; do
  echo "== pair=$pair flags=$flags k=10"; IA_RETR_PAIR=$pair IA_RETR_FLAGS=$flags timeout 120 python scripts/prof_retrieval.py 10000 1000000 1024 cosine 10 6 2>&1 | tail -n 2
done; done
for pair in 0 1; do
  echo "== pair=$pair k=100"; IA_RETR_PAIR=$pair timeout 120 python scripts/prof_retrieval.py 10000 1000000 1024 cosine 100 6 2>&1 | tail -n 2
done
for pair in 0 1; do
  IA_RETR_PAIR=$pair timeout 300 ncu --set full --clock-control none -k regex:retrieve_tc_kernel -s 1 -c 1 -o $O/ncu_retr_pair$pair -f python scripts/prof_retrieval.py 4096 500000 1024 cosine 10 2 > $O/ncu_retr_pair$pair.log 2>&1; tail -n 1 $O/ncu_retr_pair$pair.log
done
