"""Tiny driver for ncu: builds the 512^3 config-2 grid on the device and runs the SDF path."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from voxelized_geometry_tools_b200 import device as vdev, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
occupancy = synthetic.clustered_spheres_occupancy_torch((n, n, n), dev)
out = torch.empty_like(occupancy)
for _ in range(repeats):
    sdf, min_max = vdev.signed_distance_field(occupancy, 0.02, out=out)
torch.cuda.synchronize()
print("min/max", min_max.tolist())
