#!/bin/bash
# in-run A/B of retrieval tuning knobs (same box, back to back)
mkdir -p gpurun_out
for pen in 0.6 0.2 1.5 0.0; do
  for args in "10000 1000000 1024 cosine 100" "10000 1000000 1024 cosine 10" "2048 1000000 1024 cosine 100"; do
    IA_RETR_PENALTY=$pen timeout 300 python scripts/prof_retrieval.py $args 2>&1 | tail -n 2 | tr '\n' ' '; echo " [penalty $pen]"
  done
done | tee gpurun_out/retr_ab.log
for flags in 6 7 2; do
    IA_RETR_FLAGS=$flags timeout 300 python scripts/prof_retrieval.py 10000 1000000 1024 cosine 100 2>&1 | tail -n 2 | tr '\n' ' '; echo " [flags $flags]"
done | tee -a gpurun_out/retr_ab.log
