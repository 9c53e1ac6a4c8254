#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table:
launches, total and average duration and share of the step per kernel (our kernels only)."""
import csv
import re
import sys
from collections import OrderedDict

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
ours = OrderedDict()
for r in rows:
    name = r[4]
    if "vgt_b200" not in name and "unnamed>" not in name:
        continue
    short = re.sub(r"\(.*", "", name).replace("void ", "").strip()
    nanos = float(r[14])
    entry = ours.setdefault(short, [0, 0.0])
    entry[0] += 1
    entry[1] += nanos
total = sum(v[1] for v in ours.values())
print(title)
print(f"Filter: our kernels only. {sum(v[0] for v in ours.values())} launches, "
      f"{total / 1e6:.3f} ms total under ncu (cold-cache, serialised: compare shares).\n")
print("| kernel | launches | total ms | avg ms | share |")
print("|---|---|---|---|---|")
for name, (count, nanos) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {count} | {nanos / 1e6:.3f} | {nanos / 1e6 / count:.4f} | "
          f"{100 * nanos / total:.1f}% |")
