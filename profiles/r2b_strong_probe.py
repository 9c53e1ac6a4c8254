"""Why does the 1024^3 leg of bench.py take longer than the sum of its passes? Times the plain
device entry at two resolutions next to the profile entry."""
import sys, statistics
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from voxelized_geometry_tools_b200 import device as vdev, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
occupancy = synthetic.clustered_spheres_occupancy_torch((n, n, n), dev)
out = torch.empty_like(occupancy)
min_max = torch.empty(2, dtype=torch.float32, device=dev)
for resolution in (0.02, 0.01, 0.02):
    for _ in range(3):
        vdev.signed_distance_field(occupancy, resolution, out=out, min_max=min_max)
    torch.cuda.synchronize()
    times = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); vdev.signed_distance_field(occupancy, resolution, out=out, min_max=min_max); b.record()
        torch.cuda.synchronize(); times.append(a.elapsed_time(b))
    passes = [vdev.signed_distance_field_profile(occupancy, resolution, out, min_max, kernels=True) for _ in range(5)]
    med = [round(statistics.median(p[i] for p in passes), 3) for i in range(5)]
    print(n, "resolution", resolution, "plain entry ms", round(statistics.median(times), 3), "profile passes", med, "sum", round(sum(med[:3]), 3), flush=True)
