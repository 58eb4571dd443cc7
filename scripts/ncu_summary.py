#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on the CPU box) into a small JSON + markdown for profiles/.

usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01/name
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "gpu_dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_elapsed_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__cycles_elapsed.max": "sm_cycles",
    "smsp__inst_executed.sum": "warp_instructions",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "global_load_sectors",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "global_load_requests",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, vals):
            for k, name in KEYS.items():
                if h == k or h.endswith("." + k):
                    d[name] = {"value": v, "unit": u}
        kernels.append(d)
    json.dump(kernels, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full summary of `{rep}`\n\n")
        for d in kernels:
            f.write(f"## {d['kernel'][:140]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k, v in d.items():
                if k != "kernel":
                    f.write(f"| {k} | {v['value']} | {v['unit']} |\n")
            f.write("\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
