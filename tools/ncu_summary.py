#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of counters DESIGN.md cites.
usage: ncu_summary.py report.ncu-rep [> profiles/xxx.txt]"""
import csv, subprocess, sys, io
KEYS = [
 "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
 "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
 "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
 "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "lts__t_sectors.sum", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
 "l1tex__t_sectors.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
 "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
 "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
 "smsp__thread_inst_executed_per_inst_executed.ratio",
 "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]
def num(d, k):
    if k not in d:      # some counters only appear under a section prefix ("SM_B.TriageCompute.<name>")
        k = next((h for h in d if h.endswith("." + k)), k)
    try:
        return float(d[k].replace(",", ""))
    except (KeyError, ValueError):
        return None
def derived(d, u):
    """Achieved L1 / L2 / DRAM bytes per second of the launch (32-byte sectors over the kernel's duration)."""
    t = num(d, "gpu__time_duration.sum")
    if not t:
        return
    t *= {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u.get("gpu__time_duration.sum", "ns"), 1e-9)
    l1, l2 = num(d, "l1tex__t_sectors.sum"), num(d, "lts__t_sectors.sum")
    dr, dw = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
    if l1:
        print(f"  {'derived: L1 (l1tex) bytes = 32 B x sectors, per second':75s} {l1 * 32 / t / 1e9:18.1f} GB/s")
    if l2:
        print(f"  {'derived: L2 (lts) bytes = 32 B x sectors, per second':75s} {l2 * 32 / t / 1e9:18.1f} GB/s")
    if dr is not None and dw is not None:
        sc = lambda k: {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u.get(k, "byte"), 1)
        print(f"  {'derived: DRAM bytes (read + write), per second':75s} {(dr * sc('dram__bytes_read.sum') + dw * sc('dram__bytes_write.sum')) / t / 1e9:18.1f} GB/s")
    rq = num(d, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
    if l1 and rq:
        print(f"  {'derived: sectors per global load request (32 = fully divergent)':75s} {l1 / rq:18.2f}")
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
        derived(d, u)
        st = [(float(d[h]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and d[h]]
        if not st:
            st = [(float(d[h]), h) for h in hdr if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct") and d[h]]
        for v, h in sorted(st, reverse=True)[:8]:
            print(f"  stall {h:69s} {v:18.3f}")
if __name__ == "__main__":
    main()
