"""Per-step times of a 1024^3 SDF right after a burst of 512^3 calls (what bench.py does)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from voxelized_geometry_tools_b200 import device as vdev, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
small = synthetic.clustered_spheres_occupancy_torch((512, 512, 512), dev)
small_out = torch.empty_like(small)
mm = torch.empty(2, dtype=torch.float32, device=dev)
for _ in range(25):
    vdev.signed_distance_field(small, 0.02, out=small_out, min_max=mm)
for _ in range(20):
    vdev.signed_distance_field_profile(small, 0.02, small_out, mm, kernels=True)
torch.cuda.synchronize()
print("pool reserved MiB", torch.cuda.memory_reserved() >> 20, flush=True)
big = synthetic.clustered_spheres_occupancy_torch((1024, 1024, 1024), dev)
out = torch.empty_like(big)
times = []
for _ in range(14):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); vdev.signed_distance_field(big, 0.01, out=out, min_max=mm); b.record()
    torch.cuda.synchronize(); times.append(round(a.elapsed_time(b), 2))
print("per step, synchronised between steps:", times, flush=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    vdev.signed_distance_field(big, 0.01, out=out, min_max=mm)
b.record(); torch.cuda.synchronize()
print("10 steps back to back, per step:", round(a.elapsed_time(b) / 10, 2), flush=True)
free, total = torch.cuda.mem_get_info()
print("free GiB", free >> 30, "of", total >> 30)
