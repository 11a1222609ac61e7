#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of counters DESIGN.md cites.
usage: ncu_summary.py report.ncu-rep [> profiles/xxx.txt]"""
import csv, subprocess, sys, io
KEYS = [
 "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
 "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
 "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
 "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
 "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
 "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
 "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
 "smsp__thread_inst_executed_per_inst_executed.ratio",
 "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]
def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print(f"kernel: {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]:>18s} {u[k]}")
        st = [(float(d[h]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and d[h]]
        if not st:
            st = [(float(d[h]), h) for h in hdr if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct") and d[h]]
        for v, h in sorted(st, reverse=True)[:8]:
            print(f"  stall {h:69s} {v:18.3f}")
if __name__ == "__main__":
    main()
