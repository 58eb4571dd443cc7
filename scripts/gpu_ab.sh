#!/bin/bash
# in-run A/B of retrieval tuning flags (same box, back to back, two rounds)
mkdir -p gpurun_out
for round in 1 2; do
for flags in 7 6 5 3 0; do
  for args in "10000 1000000 1024 cosine 100" "10000 1000000 1024 cosine 10"; do
    IA_RETR_FLAGS=$flags timeout 300 python scripts/prof_retrieval.py $args 2>&1 | tail -n 3 | tr '\n' ' '; echo
  done
done
done | tee gpurun_out/retr_ab.log
