#!/usr/bin/env python3
"""Distil one `ncu --set full` capture of the trace kernel into the numbers bench.py quotes next to its roofline
(profiles/roofline_evidence.json): DRAM bytes per launch, issue-slot and lane utilisation, achieved L1 / L2 sector
rates against the ceilings of the kernel's access pattern measured by tools/gather_ceiling.cu.
usage: ncu_evidence.py <workload> <report.ncu-rep> <label of the capture> [gather_ceiling.json]   (run here, no GPU)"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
workload, rep, label = sys.argv[1], sys.argv[2], sys.argv[3]
ceil_path = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "r02_gather_ceiling.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
d, u = dict(zip(hdr, r)), dict(zip(hdr, units))


def num(k):
    if k not in d:
        k = next((h for h in d if h.endswith("." + k)), k)
    return float(d[k].replace(",", ""))


def scaled(k):
    return num(k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u.get(k, "byte"), 1)


t = num("gpu__time_duration.sum") * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[u["gpu__time_duration.sum"]]
l1s, l2s = num("l1tex__t_sectors.sum"), num("lts__t_sectors.sum")
req = num("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
spr = l1s / req
ceil = json.load(open(ceil_path))["results"]


def ceiling(level):
    """sector GB/s of the microbenchmark at the nearest measured sectors-per-request"""
    c = min((x for x in ceil if x["level"] == level), key=lambda x: abs(x["sectors_per_request"] - spr))
    return c


c1, c2 = ceiling("L1"), ceiling("L2")
l1_gbs, l2_gbs = l1s * 32 / t / 1e9, l2s * 32 / t / 1e9
issue = num("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100
lanes = num("smsp__thread_inst_executed_per_inst_executed.ratio")
dram = scaled("dram__bytes_read.sum") + scaled("dram__bytes_write.sum")
dram_pct = num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
ev = {
    "source": f"{label}; ceilings: {os.path.relpath(ceil_path, ROOT)} (tools/gather_ceiling.cu)",
    "kernel": d.get("Kernel Name"), "kernel_ms_under_ncu": t * 1e3,
    "dram_bytes_per_launch": dram, "dram_busy_pct": dram_pct,
    "issue_active": issue, "lanes_active": lanes,
    "warp_instructions_per_launch": num("smsp__inst_executed.sum"),
    "l1": {"achieved_GBps": l1_gbs, "sectors_per_request": spr, "hit_rate": num("l1tex__t_sector_hit_rate.pct") / 100,
           "requests_per_s": req / t,
           "ceiling_GBps": c1["sector_GBps"], "ceiling_requests_per_s": c1["requests_per_s"],
           "ceiling_at_sectors_per_request": c1["sectors_per_request"], "frac": l1_gbs / c1["sector_GBps"],
           "lsu_data_pipe_busy": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100},
    "l2": {"achieved_GBps": l2_gbs, "hit_rate": num("lts__t_sector_hit_rate.pct") / 100,
           "ceiling_GBps": c2["sector_GBps"], "ceiling_at_sectors_per_request": c2["sectors_per_request"],
           "frac": l2_gbs / c2["sector_GBps"]},
    "limiter": (f"instruction issue: {issue * 100:.0f} % of issue slots active at {lanes:.1f} of 32 lanes; the working set is "
                f"L2-resident (DRAM {dram_pct:.2f} % busy, {dram / 1e6:.0f} MB per launch), L1 at {l1_gbs / c1['sector_GBps'] * 100:.0f} % and "
                f"L2 at {l2_gbs / c2['sector_GBps'] * 100:.0f} % of the measured gather ceilings"),
}
out = os.path.join(ROOT, "profiles", "roofline_evidence.json")
try:
    allv = json.load(open(out))
except Exception:
    allv = {}
allv[workload] = ev
json.dump(allv, open(out, "w"), indent=1)
print(json.dumps(ev, indent=1))
