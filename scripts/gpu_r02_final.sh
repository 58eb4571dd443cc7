#!/bin/bash
# round 2, final single-GPU evidence after the last kernel changes
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -n 2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 1
timeout 900 python bench.py > $O/bench_n1_final_r02.json 2> $O/bench_n1_final_r02.err; echo "bench rc $?"; tail -c 1500 $O/bench_n1_final_r02.json; echo
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference_final_r02.json 2>/dev/null; echo "reference rc $?"
timeout 200 python scripts/bench_pair_configs.py > $O/pair_configs_final_r02.log 2>&1; tail -n 10 $O/pair_configs_final_r02.log
timeout 200 python scripts/prof_projection_train.py 6 2>/dev/null
timeout 300 python - <<'PY' 2>&1 | tail -n 3
import sys, time; sys.path.insert(0, '.')
import torch, item_alignment_b200 as ia
for dt in (torch.float32, torch.float64):
    s = torch.sigmoid(torch.randn(10_000_000, device='cuda')).to(torch.bfloat16).to(dt); l = (torch.rand(10_000_000, device='cuda') < 0.3).long()
    for _ in range(2): ia.find_best_f1_and_threshold(s, l)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): r = ia.find_best_f1_and_threshold(s, l)
    print(f"best-F1 threshold search, 10^7 {dt} scores: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per call (kernels + 40-byte read-back)", r[1], r[4])
    import oracle  # noqa
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 5 -c 1 -o $O/ncu_pair_final_r02 -f python scripts/exp_pair_gap.py > $O/ncu_pair_final_r02.log 2>&1; tail -n 1 $O/ncu_pair_final_r02.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/ncu_launch_list_bench_final_r02.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-config5 > /dev/null 2>&1; grep -c "ia::" $O/ncu_launch_list_bench_final_r02.csv
