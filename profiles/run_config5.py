"""BASELINE config 5 under torchrun on 8 GPUs: 2048^3 clustered spheres and the 4096x4096x512
terrain, slab-sharded with the fused NVLink exchange. Checks: (1) sign == class and |d| >= one
voxel on every slab, (2) the sharded result equals the single-GPU result slab by slab (exact
checksums of the float bit patterns; rank 0 recomputes the whole grid on its own GPU), then times
the sharded path. Prints one JSON line per grid on rank 0."""
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from voxelized_geometry_tools_b200 import device as vdev, synthetic  # noqa: E402
from voxelized_geometry_tools_b200.sharded import ShardedSignedDistanceField, split_range  # noqa: E402

RESOLUTION = 0.01


def checksum(sdf: torch.Tensor):
    bits = sdf.contiguous().view(torch.int32).to(torch.int64)
    weights = torch.arange(bits.numel(), device=bits.device, dtype=torch.int64).view_as(bits) % 1021
    return int(bits.sum().item()), int((bits * (weights + 1)).sum().item())


def main():
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    scale = int(os.environ.get("CONFIG5_SCALE", "1"))   # 2 = half-size dry run
    grids = [("2048^3 clustered spheres", (2048 // scale,) * 3,
              lambda dims, x_range: synthetic.clustered_spheres_occupancy_torch(
                  dims, dev, x_range=x_range)),
             ("4096x4096x512 terrain", (4096 // scale, 4096 // scale, 512 // scale),
              lambda dims, x_range: synthetic.terrain_occupancy_torch(dims, dev, x_range=x_range))]
    for name, dims, generate in grids:
        plan = ShardedSignedDistanceField(dims)
        slab = generate(dims, plan.x_range)
        sdf, min_max = plan.extract(slab, RESOLUTION)
        torch.cuda.synchronize(dev)
        # (1) per-slab properties: the y-slab of the occupancy vs the sign of the SDF
        y0, y1 = plan.y_range
        ok = True
        for peer in range(world):          # rebuild my y-slab's classes from every x-slab
            x0, x1 = split_range(dims[0], world, peer)
            part = generate(dims, (x0, x1))[:, y0:y1, :]
            filled = part >= 0.5
            piece = sdf[x0:x1]
            ok &= bool(torch.all(piece[filled] < 0)) and bool(torch.all(piece[~filled] > 0))
            ok &= bool(piece.abs().min() >= RESOLUTION * (1 - 1e-6))
            del part, filled
        mine = checksum(sdf)
        gathered = [None] * world
        dist.all_gather_object(gathered, (ok, mine))
        # timing of the sharded path
        for _ in range(2):
            plan.extract(slab, RESOLUTION)
        dist.barrier()
        torch.cuda.synchronize(dev)
        begin = time.perf_counter()
        steps = 5
        for _ in range(steps):
            plan.extract(slab, RESOLUTION)
        torch.cuda.synchronize(dev)
        dist.barrier()
        seconds = (time.perf_counter() - begin) / steps
        del slab, sdf
        torch.cuda.empty_cache()
        if rank == 0:
            # (2) single-GPU recomputation of the whole grid, compared slab by slab
            whole = generate(dims, None)
            single, single_min_max = vdev.signed_distance_field(whole, RESOLUTION)
            del whole
            same = True
            for peer in range(world):
                py0, py1 = split_range(dims[1], world, peer)
                same &= checksum(single[:, py0:py1, :]) == gathered[peer][1]
            same &= single_min_max.tolist() == min_max.tolist()
            del single
            torch.cuda.empty_cache()
            voxels = dims[0] * dims[1] * dims[2]
            print(json.dumps({
                "grid": name, "dims": dims, "n_gpus": world, "exchange": plan.exchange_used,
                "properties_ok": all(g[0] for g in gathered), "equals_single_gpu": bool(same),
                "ms_per_sdf": seconds * 1e3, "gvoxels_per_s": voxels / seconds / 1e9,
                "hbm_roofline_frac_24B": 24 * voxels / seconds / 1e9 / (6541.1 * world),
                "sdf_min_max": min_max.tolist()}), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
