"""Times the three SDF passes on one B200 (device-resident, CUDA events inside the C-ABI profile
entry) for a clustered-spheres grid. Usage: time_passes.py [n] [repeats]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from voxelized_geometry_tools_b200 import device as vdev, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
occupancy = synthetic.clustered_spheres_occupancy_torch((n, n, n), dev)
out = torch.empty_like(occupancy)
for _ in range(3):
    vdev.signed_distance_field(occupancy, 0.02, out=out)
torch.cuda.synchronize()
rows = []
for _ in range(repeats):
    rows.append(vdev.signed_distance_field_profile(occupancy, 0.02, out=out))
torch.cuda.synchronize()
rows = torch.tensor(rows)
med = rows.median(dim=0).values.tolist()
print("n", n, "median ms z/y/x", [round(v, 4) for v in med], "total", round(sum(med), 4))
