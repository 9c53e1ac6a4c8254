#!/usr/bin/env python3
"""Experiment: y-pass window radius (VGT_B200_WINDOW_RADIUS_Y = 12 / 10 / 8): pass times at 512^3
and 1024^3, and that the SDF does not change."""
import os
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from voxelized_geometry_tools_b200 import _capi, device as vdev, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
for n in (512, 1024):
    occupancy = synthetic.clustered_spheres_occupancy_torch((n, n, n), dev)
    reference = None
    for radius, stage in (("12", ""), ("12", "1"), ("12", "2"), ("12", "0")):
        os.environ["VGT_B200_WINDOW_RADIUS_Y"] = radius
        os.environ.pop("VGT_B200_WINDOW_STAGE", None)
        if stage:
            os.environ["VGT_B200_WINDOW_STAGE"] = stage
        _capi.library().vgt_b200_reload_tuning()
        out = torch.empty_like(occupancy)
        min_max = torch.empty(2, device=dev)
        for _ in range(2):
            vdev.signed_distance_field(occupancy, 0.02, out=out, min_max=min_max)
        samples = [vdev.signed_distance_field_profile(occupancy, 0.02, out, min_max)
                   for _ in range(7)]
        passes = [round(statistics.median(s[i] for s in samples), 4) for i in range(3)]
        if reference is None:
            reference = out.clone()
        print(n, "radius", radius, "stage", stage or "default", "passes", passes, "total", round(sum(passes), 4),
              "same", bool(torch.equal(out, reference)), flush=True)
    del occupancy, reference, out
    torch.cuda.empty_cache()
