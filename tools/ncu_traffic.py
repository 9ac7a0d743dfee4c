#!/usr/bin/env python3
"""DRAM traffic per launch of every kernel instance in an .ncu-rep (dram__bytes_read.sum + dram__bytes_write.sum) as JSON."""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
launches = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))

    def to_bytes(key):
        v = float(d[key])
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[key]]

    launches.append({"kernel": d["Kernel Name"].split("(")[0], "duration_ms_under_ncu": float(d["gpu__time_duration.sum"]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u["gpu__time_duration.sum"], 1.0),
                     "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum")})
tot = [x["dram_bytes_read"] + x["dram_bytes_write"] for x in launches]
print(json.dumps({"source": rep.split("/")[-1], "command": " ".join(sys.argv[2:]), "launches": launches,
                  "dram_bytes_per_launch_mean": sum(tot) / max(len(tot), 1)}, indent=1))
