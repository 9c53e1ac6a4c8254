"""Launched under torchrun (one rank per GPU): the NCCL slab-sharded SDF must equal the 1-GPU
result bit for bit, and the oracle at a size the oracle finishes quickly."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from voxelized_geometry_tools_b200 import device as vdev, synthetic  # noqa: E402
from voxelized_geometry_tools_b200.sharded import ShardedSignedDistanceField  # noqa: E402


def check_sharded_voxelizer(dev, rank, world):
    """Rays split over the ranks + reduce-scatter of the counters == the one-GPU voxelizer, then
    straight into the sharded SDF == the one-GPU SDF of the one-GPU map."""
    from voxelized_geometry_tools_b200.grids import compose_rigid, inverse_rigid
    from voxelized_geometry_tools_b200.pointcloud_voxelization import (
        PointCloudVoxelizationFilterOptions)
    from voxelized_geometry_tools_b200.sharded import ShardedPointCloudVoxelizer
    for grid_n in (64, 100):      # 100: x-slabs of uneven size at world 8
        scene = synthetic.depth_camera_scene(grid_n, 5.12 / grid_n, 160, 120, max_range=5.0)
        dims = scene["static_occupancy"].shape
        x_gw = inverse_rigid(scene["origin_transform"])
        clouds = [(torch.from_numpy(p).to(dev), compose_rigid(x_gw, x), r)
                  for p, x, r in scene["clouds"]]
        options = PointCloudVoxelizationFilterOptions(0.9, 2, 2)
        voxelizer = ShardedPointCloudVoxelizer(dims, scene["voxel_size"])
        x0, x1 = voxelizer.x_range
        static = torch.from_numpy(scene["static_occupancy"]).to(dev)
        slab = voxelizer.voxelize(static[x0:x1].contiguous(), clouds, options, keep_counts=True)
        # one GPU, all rays
        counts = torch.zeros((len(clouds),) + tuple(dims) + (2,), dtype=torch.int32, device=dev)
        for index, (points, x_gc, max_range) in enumerate(clouds):
            vdev.raycast_cloud(points, x_gc, max_range, counts[index], scene["voxel_size"])
        whole = static.clone()
        vdev.filter_grids(counts, whole, options)
        assert torch.equal(voxelizer.last_counts, counts[:, x0:x1]), "sharded counts differ"
        assert torch.equal(slab, whole[x0:x1]), "sharded occupancy differs"
        plan = ShardedSignedDistanceField(dims)
        sdf_slab, _ = plan.extract(slab, scene["voxel_size"])
        full = plan.gather_to_host(sdf_slab)
        if rank == 0:
            single, _ = vdev.signed_distance_field(whole, scene["voxel_size"])
            assert torch.equal(full, single.cpu()), "voxelize -> SDF: sharded != single GPU"
            print(f"sharded voxelizer {grid_n}^3 -> SDF: world {world} ok", flush=True)
        dist.barrier()


def main():
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    for dims, border, chunks, exchange in (
            ((96, 80, 72), False, 1, "nccl"), ((61, 45, 130), True, 3, "nccl"),
            ((256, 256, 256), False, 4, "nccl"), ((96, 80, 72), False, 1, "peer_store"),
            ((61, 45, 130), True, 1, "peer_store"), ((256, 256, 256), False, 1, "peer_store"),
            ((96, 80, 72), False, 2, "peer_copy"), ((61, 45, 130), True, 3, "peer_copy"),
            ((256, 256, 256), False, 4, "peer_copy")):
        plan = ShardedSignedDistanceField(dims, chunks=chunks, exchange=exchange)
        slab = synthetic.clustered_spheres_occupancy_torch(dims, dev, x_range=plan.x_range)
        for _ in range(3):      # several steps: exercises the double-buffered receive side
            sdf_slab, min_max = plan.extract(slab, 0.02, add_virtual_border=border)
        assert plan.exchange_used == exchange, (plan.exchange_used, exchange)
        full = plan.gather_to_host(sdf_slab)
        if rank == 0:
            whole = synthetic.clustered_spheres_occupancy_torch(dims, dev)
            single, single_min_max = vdev.signed_distance_field(whole, 0.02,
                                                                add_virtual_border=border)
            assert torch.equal(full, single.cpu()), f"{dims}: sharded != single GPU"
            assert min_max.tolist() == single_min_max.tolist()
            if np.prod(dims) <= 2 ** 21:
                from oracle import oracle
                want, _ = oracle.sdf(whole.cpu().numpy(), 0.02, add_virtual_border=border)
                assert np.array_equal(full.numpy(), want), f"{dims}: sharded != oracle"
            print(f"dims {dims} border {border} {exchange}: world {world} ok", flush=True)
        dist.barrier()
    check_sharded_voxelizer(dev, rank, world)
    if rank == 0:
        print("MULTI_GPU_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
