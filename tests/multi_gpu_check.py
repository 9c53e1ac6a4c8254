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


def main():
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    for dims, border, chunks, exchange in (
            ((96, 80, 72), False, 1, "nccl"), ((61, 45, 130), True, 3, "nccl"),
            ((256, 256, 256), False, 4, "nccl"), ((96, 80, 72), False, 1, "peer_store"),
            ((61, 45, 130), True, 1, "peer_store"), ((256, 256, 256), False, 1, "peer_store")):
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
    if rank == 0:
        print("MULTI_GPU_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
