#!/usr/bin/env python
"""Development tool: write profiles/traffic.json from an `ncu --set full` capture of bench.py's sweep launches (cavity 512^3).
usage: tools/ncu_traffic.py report.ncu-rep "where the capture is summarised"
The file carries a hash of the kernel sources; bench.py reports `roofline.traffic` only while that hash matches."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_sources_hash  # noqa: E402

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
name, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


per = [to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]) for r in data if "sweep_kernel" in r[name]]
assert per, "no sweep_kernel launch in the report"
json.dump({"dram_bytes_per_launch_512": sum(per) / len(per), "launches": len(per), "per_launch": per,
           "kernel_sources_sha256": kernel_sources_hash(), "source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", "traffic.json")).read())
