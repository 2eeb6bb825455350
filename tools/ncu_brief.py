#!/usr/bin/env python3
"""tools/ncu_brief.py <report.ncu-rep> : per-kernel digest (time, instructions, issue, stalls, smem, dram)."""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def g(r, k):
    try: return float(r[hdr.index(k)].replace(",", ""))
    except Exception: return float("nan")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")][:50]
    print("==== %s  %.1f %s" % (name, g(r, "gpu__time_duration.sum"), units[hdr.index("gpu__time_duration.sum")]))
    wi = g(r, "smsp__inst_executed.sum"); lanes = g(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
    print("  warp-inst %.1fM  lanes/inst %.1f  thread-inst %.2fG  issue-active %.1f%%  warps/sched %.1f eligible %.2f" % (
        wi / 1e6, lanes, wi * lanes / 1e9, g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        g(r, "smsp__warps_active.avg.per_cycle_active"), g(r, "smsp__warps_eligible.avg.per_cycle_active")))
    print("  pipes alu %.0f fma %.0f lsu %.0f | regs %d  occ-limit smem %d regs %d blocks" % (
        g(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), g(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        g(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), g(r, "launch__registers_per_thread"),
        g(r, "launch__occupancy_limit_shared_mem"), g(r, "launch__occupancy_limit_registers")))
    print("  smem wavefronts %.1fM (conflicts %.1fM; ld %.1fM st %.1fM)  global ld sectors %.1fM req %.2fM  L1 hit %.0f%%" % (
        g(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / 1e6, g(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / 1e6,
        g(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum") / 1e6, g(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum") / 1e6,
        g(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum") / 1e6, g(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum") / 1e6,
        g(r, "l1tex__t_sector_hit_rate.pct")))
    print("  dram rd %.1f %s wr %.1f %s" % (g(r, "dram__bytes_read.sum"), units[hdr.index("dram__bytes_read.sum")],
                                          g(r, "dram__bytes_write.sum"), units[hdr.index("dram__bytes_write.sum")]))
    st = [(g(r, h), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
          for h in hdr if h.startswith("smsp__average_warps_issue_stalled_")]
    print("  stalls/issue: " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:7]))
