#!/usr/bin/env python3
"""One rank's share of BASELINE config 5 (2048^3 on 8 GPUs) on ONE device, no exchange: the
slab-local passes of x-slab 0 and the final pass of y-slab 0 (its input taken from the local
passes of 2048 x 256 x 2048 worth of x-planes is not available on one device, so the final pass
runs on the y-slab of a 1024-plane half grid: same line length along y and z, half along x)."""
import json
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from voxelized_geometry_tools_b200 import device as vdev, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
dims = (2048, 2048, 2048)


def timed(fn, reps=3):
    times = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return statistics.median(times)


import os
from voxelized_geometry_tools_b200 import _capi
result = {"dims": list(dims)}
slab = synthetic.clustered_spheres_occupancy_torch(dims, dev, x_range=(0, 256))
for name, env in (("default", {}), ("budget100000", {"VGT_B200_WINDOW_BUDGET": "100000"}),
                  ("lean", {"VGT_B200_ENVELOPE": "lean"}), ("nopilot", {"VGT_B200_WINDOW_PILOT": "0"})):
    for key in ("VGT_B200_WINDOW_BUDGET", "VGT_B200_ENVELOPE", "VGT_B200_WINDOW_PILOT"):
        os.environ.pop(key, None)
    os.environ.update(env)
    _capi.library().vgt_b200_reload_tuning()
    vdev.edt_local_passes(slab, send_parts=8)
    result[f"local_passes_ms_{name}"] = timed(lambda: vdev.edt_local_passes(slab, send_parts=8))
print(json.dumps(result), flush=True)
