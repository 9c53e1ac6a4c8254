#!/usr/bin/env python3
"""Experiment (one GPU): the 2048^3 clustered-spheres grid of BASELINE config 5 on a single B200
(103 GB of buffers), pass by pass, under different step allowances of the window kernel's joint
search (VGT_B200_WINDOW_BUDGET, percent: search steps per 100 rows before a tile is handed to the
stack kernel) and with the stack kernel alone. Shows where the time of deep maps goes.

    python profiles/r2_deep_2048.py [n] [budgets ...]
"""
import os
import statistics
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from voxelized_geometry_tools_b200 import _capi, device as vdev, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
budgets = sys.argv[2:] or ["2400", "800", "300", "100", "lean"]
dev = torch.device("cuda", 0)
t0 = time.time()
occupancy = synthetic.clustered_spheres_occupancy_torch((n, n, n), dev)
torch.cuda.synchronize()
torch.cuda.empty_cache()
print("generated", n, "in", round(time.time() - t0, 1), "s", flush=True)
out = torch.empty_like(occupancy)
min_max = torch.empty(2, device=dev)
reference = None
for budget in budgets:
    os.environ.pop("VGT_B200_ENVELOPE", None)
    os.environ.pop("VGT_B200_WINDOW_BUDGET", None)
    if budget == "lean":
        os.environ["VGT_B200_ENVELOPE"] = "lean"
    else:
        os.environ["VGT_B200_WINDOW_BUDGET"] = budget
    _capi.library().vgt_b200_reload_tuning()
    vdev.signed_distance_field(occupancy, 0.02, out=out, min_max=min_max)
    samples = [vdev.signed_distance_field_profile(occupancy, 0.02, out, min_max) for _ in range(3)]
    passes = [round(statistics.median(s[i] for s in samples), 3) for i in range(3)]
    checksum = (float(out.double().sum().item()), min_max.tolist())
    if reference is None:
        reference = checksum
    print(n, "budget", budget, "passes z/y/x", passes, "total", round(sum(passes), 3),
          "Gvoxels/s", round(n ** 3 / sum(passes) / 1e6, 1), "same", checksum == reference, flush=True)
