#!/usr/bin/env python3
"""Writes profiles/ncu_traffic.json (per-launch DRAM bytes of the three SDF kernels) from an
`ncu --set full` report of `profiles/run_sdf_once.py <n>`: ncu_traffic.py <report.ncu-rep> <n>"""
import csv
import json
import subprocess
import sys
from pathlib import Path

report, n = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
header, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def value(row, name):
    i = header.index(name)
    return float(row[i]) * scale[units[i]]


labels = {"ScanContiguousAxis": "ScanContiguousAxisRegistersKernel (z)",
          "EnvelopeAxisWindowKernel<0": "EnvelopeAxisWindowKernel (y)",
          "EnvelopeAxisWindowKernel<1": "EnvelopeAxisWindowKernel (x + finalize)"}
out = {}
for row in rows[2:]:
    name = row[header.index("Kernel Name")]
    for key, label in labels.items():
        # (the window kernels launch twice per pass: the short pilot probe first, then the pass
        # itself; keep the longer launch)
        duration = float(row[header.index("gpu__time_duration.sum")]) \
            * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[units[header.index("gpu__time_duration.sum")]]
        if key in name and (label not in out or duration > out[label]["gpu_time_ms_under_ncu"]):
            out[label] = {"dram_bytes_read": value(row, "dram__bytes_read.sum"),
                          "dram_bytes_write": value(row, "dram__bytes_write.sum"),
                          "gpu_time_ms_under_ncu": float(row[header.index("gpu__time_duration.sum")])
                          * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[units[header.index("gpu__time_duration.sum")]],
                          "report": Path(report).name}
path = Path(__file__).resolve().parent / "ncu_traffic.json"
table = json.loads(path.read_text()) if path.exists() else {}
# The capture describes ONE build of the library: bench.py attaches these numbers only while the
# kernel sources still hash to the same value (run this right after the capture, before editing).
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from voxelized_geometry_tools_b200 import build as cuda_build  # noqa: E402
for entry in out.values():
    entry["sources_sha1"] = cuda_build._sources_signature()
table[f"{n}x{n}x{n}"] = out
path.write_text(json.dumps(table, indent=1) + "\n")
print(json.dumps(out, indent=1))
