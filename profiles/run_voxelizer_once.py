"""Tiny driver for ncu: BASELINE config 3 (4 cameras x 640x480 rays into 256^3) raycast + filter
on device-resident data."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from voxelized_geometry_tools_b200 import device as vdev, synthetic  # noqa: E402
from voxelized_geometry_tools_b200.grids import compose_rigid  # noqa: E402
from voxelized_geometry_tools_b200.pointcloud_voxelization import (  # noqa: E402
    PointCloudVoxelizationFilterOptions)

dev = torch.device("cuda", 0)
scene = synthetic.depth_camera_scene()
n = scene["static_occupancy"].shape[0]
x_gw = np.eye(4)
x_gw[:3, 3] = -scene["origin_transform"][:3, 3]
counts = torch.zeros((len(scene["clouds"]), n, n, n, 2), dtype=torch.int32, device=dev)
occupancy = torch.from_numpy(scene["static_occupancy"]).to(dev)
for index, (points, x_wc, max_range) in enumerate(scene["clouds"]):
    vdev.raycast_cloud(torch.from_numpy(points).to(dev), compose_rigid(x_gw, x_wc), max_range,
                       counts[index], scene["voxel_size"])
vdev.filter_grids(counts, occupancy, PointCloudVoxelizationFilterOptions(0.9, 2, 2))
torch.cuda.synchronize()
print("increments", int(counts.sum(dtype=torch.int64).item()))
