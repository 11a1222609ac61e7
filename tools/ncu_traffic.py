#!/usr/bin/env python3
"""Write profiles/roofline_traffic.json from ncu --set full captures:
   ncu_traffic.py workload=report.ncu-rep [workload=report.ncu-rep ...]"""
import csv, io, json, os, subprocess, sys
out = {}
for spec in sys.argv[1:]:
    wl, rep = spec.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
    wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
    out[wl] = {"kernel": d["Kernel Name"], "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
               "gpu_time_ms": float(d["gpu__time_duration.sum"]) * {"ms": 1, "us": 1e-3, "ns": 1e-6}.get(u["gpu__time_duration.sum"], 1),
               "source": os.path.basename(rep)}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, "profiles", "roofline_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
