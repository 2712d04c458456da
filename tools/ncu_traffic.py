#!/usr/bin/env python
"""profiles/<rep>.ncu-rep -> profiles/k_match_fast_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of the
k_match_fast launches in an `ncu --set full` capture (mean per launch).  bench.py reports it as roofline.traffic."""
import csv
import json
import subprocess
import sys
from pathlib import Path

rep = Path(sys.argv[1])
kernel = sys.argv[2] if len(sys.argv) > 2 else "k_match_fast"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ki, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, n, ms = 0.0, 0, 0.0
for r in rows[2:]:
    if kernel in r[ki]:
        tot += float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
        ms += float(r[ti]) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(units[ti], 1.0)
        n += 1
assert n, f"no {kernel} launch in {rep}"
out = {"dram_bytes_per_launch": tot / n, "launches": n, "kernel": kernel, "ncu_ms_per_launch": ms / n,
       "source": f"profiles/{rep.name} (ncu --set full --clock-control none, mean of {n} launches of {kernel})"}
dst = rep.parent / f"{kernel}_traffic.json"
dst.write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out))
