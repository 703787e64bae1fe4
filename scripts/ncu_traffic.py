#!/usr/bin/env python
"""DRAM traffic of the first kernel in an .ncu-rep (one `ncu --set full` capture) -> profiles/roofline_traffic.json,
which bench.py quotes as roofline.traffic (bytes per launch) when workload and ef match.
usage: python scripts/ncu_traffic.py report.ncu-rep WORKLOAD EF SOURCE_TAG"""
import csv, json, subprocess, sys
rep, workload, ef, tag = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
def val(k):
    i = h.index(k)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[i]]
    return float(v[i].replace(",", "")) * scale
out = {"kernel": v[h.index("Kernel Name")], "workload": workload, "ef": ef,
       "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
       "source": tag}
out["dram_bytes_per_launch"] = out["dram_bytes_read"] + out["dram_bytes_write"]
json.dump(out, open("profiles/roofline_traffic.json", "w"), indent=1)
print(out)
