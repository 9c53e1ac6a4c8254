"""Times the three SDF passes for a list of grid shapes (device-resident, C-ABI profile entry).
Usage: time_shapes.py nx,ny,nz [nx,ny,nz ...]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from voxelized_geometry_tools_b200 import device as vdev, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
for spec in sys.argv[1:]:
    dims = tuple(int(v) for v in spec.split(","))
    occupancy = synthetic.clustered_spheres_occupancy_torch(dims, dev)
    out = torch.empty_like(occupancy)
    for _ in range(2):
        vdev.signed_distance_field(occupancy, 0.02, out=out)
    torch.cuda.synchronize()
    rows = [vdev.signed_distance_field_profile(occupancy, 0.02, out=out) for _ in range(7)]
    med = torch.tensor(rows).median(dim=0).values.tolist()
    print(dims, "ms z/y/x", [round(v, 3) for v in med], "total", round(sum(med), 3), "Gvox/s",
          round(occupancy.numel() / sum(med) / 1e6, 1), flush=True)
    del occupancy, out
    torch.cuda.empty_cache()
