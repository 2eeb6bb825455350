#!/usr/bin/env python3
"""Turns .ncu-rep captures (gpurun_out/) into the small text/JSON summaries kept under profiles/.

    python tools/ncu_summary.py <name>=<file.ncu-rep> ... --out profiles/r1_kernels.json

For every kernel instance in a report: duration, DRAM bytes, issue utilisation, pipe utilisation,
shared-memory wavefronts / bank conflicts, occupancy limiters, registers, top stall reasons.
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
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_instruction",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct_of_peak",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_blocks",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
}
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def summarize(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        rec = {"kernel": None, "stalls_per_issue": {}}
        for h, u, v in zip(header, units, vals):
            if h == "Kernel Name":
                rec["kernel"] = v
            elif h in KEYS:
                try:
                    rec[KEYS[h]] = float(v)
                except ValueError:
                    rec[KEYS[h]] = v
                if u:
                    rec[KEYS[h] + "_unit"] = u
            elif h.startswith(STALL_PREFIX) and h.endswith("_per_issue_active.ratio"):
                try:
                    rec["stalls_per_issue"][h[len(STALL_PREFIX):-len("_per_issue_active.ratio")]] = round(float(v), 3)
                except ValueError:
                    pass
        rec["stalls_per_issue"] = dict(sorted(rec["stalls_per_issue"].items(), key=lambda kv: -kv[1])[:6])
        out.append(rec)
    return out


def main():
    args = sys.argv[1:]
    out_path = None
    if "--out" in args:
        i = args.index("--out")
        out_path = args[i + 1]
        del args[i:i + 2]
    result = {}
    for a in args:
        name, path = a.split("=", 1)
        result[name] = summarize(path)
    text = json.dumps(result, indent=1)
    if out_path:
        open(out_path, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
