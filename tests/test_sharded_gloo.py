"""CPU: the slab-sharded SDF driver over gloo, world_size 2 and 3.

The CUDA stages are replaced by CPU stand-ins built on the oracle's in-place transform (test
infrastructure), so this exercises exactly the host-side logic of the N > 1 path: slab ranges,
the all-to-all transpose into y-slabs, min/max reduction and the gather."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

NONE = 0x7FFFFFFF


def _cpu_stages(oracle):
    from voxelized_geometry_tools_b200.sharded import Stages

    def pack(filled, to_filled, to_free):
        squared = np.where(filled, to_free, to_filled)
        words = np.where(np.isinf(squared), NONE, squared).astype(np.int64)
        return (words | (filled.astype(np.int64) << 31)).astype(np.uint32).view(np.int32)

    def unpack(words):
        raw = words.view(np.uint32).astype(np.int64)
        filled = (raw >> 31) != 0
        value = (raw & NONE).astype(np.float64)
        value[value == NONE] = np.inf
        return filled, np.where(filled, np.inf, value), np.where(filled, value, np.inf)

    def local_passes(occupancy_slab, unknown_is_filled):
        occ = occupancy_slab.numpy()
        filled = (occ > 0.5) | (unknown_is_filled & (occ == 0.5))
        to_filled = np.where(filled, 0.0, np.inf)
        to_free = np.where(filled, np.inf, 0.0)
        for x in range(occ.shape[0]):      # per x-plane: the X pass is skipped for nx == 1
            for field in (to_filled, to_free):
                plane = np.ascontiguousarray(field[x:x + 1])
                oracle.transform_inplace(plane)
                field[x:x + 1] = plane
        return torch.from_numpy(pack(filled, to_filled, to_free))

    def final_pass(packed, y_offset, ny_total, resolution, add_virtual_border):
        assert not add_virtual_border
        filled, to_filled, to_free = unpack(packed.numpy())
        nx, nyl, nz = filled.shape
        for field, is_site in ((to_filled, filled), (to_free, ~filled)):
            field[is_site] = 0.0
            columns = np.ascontiguousarray(field.reshape(nx, nyl * nz).T)   # [columns, nx]
            for column in columns:                                         # z-only grid [1, 1, nx]
                line = np.ascontiguousarray(column.reshape(1, 1, nx))
                oracle.transform_inplace(line)
                column[:] = line.reshape(nx)
            field[...] = columns.T.reshape(nx, nyl, nz)
        magnitude = np.sqrt(np.where(filled, to_free, to_filled)) * resolution
        sdf = np.where(filled, -magnitude, magnitude).astype(np.float32)
        return torch.from_numpy(sdf), torch.tensor([sdf.min(), sdf.max()])

    def local_passes_send(occupancy_slab, unknown_is_filled, parts):
        # send layout: part h of every y-line as the block [nxl, rows_h, nz], blocks back to back
        from voxelized_geometry_tools_b200.sharded import split_range
        packed = local_passes(occupancy_slab, unknown_is_filled)
        ny = packed.shape[1]
        return torch.cat([packed[:, split_range(ny, parts, h)[0]:split_range(ny, parts, h)[1], :]
                          .reshape(-1) for h in range(parts)])

    return Stages(local_passes, final_pass, local_passes_send)


def _worker(rank, world, port, shape, result_path, chunks=1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        from voxelized_geometry_tools_b200.sharded import ShardedSignedDistanceField
        rng = np.random.default_rng(77)
        occupancy = (rng.random(shape) < 0.15).astype(np.float32)
        occupancy[rng.random(shape) < 0.05] = 0.5
        plan = ShardedSignedDistanceField(shape, stages=_cpu_stages(oracle), chunks=chunks)
        x0, x1 = plan.x_range
        sdf_slab, min_max = plan.extract(torch.from_numpy(occupancy[x0:x1].copy()), 0.25)
        assert tuple(sdf_slab.shape) == plan.y_slab_shape()
        full = plan.gather_to_host(sdf_slab)
        if rank == 0:
            want, (lo, hi) = oracle.sdf(occupancy, 0.25)
            np.testing.assert_array_equal(full.numpy(), want)
            assert float(min_max[0]) == lo and float(min_max[1]) == hi
            Path(result_path).write_text("ok")
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,shape,chunks", [(2, (10, 9, 7), 1), (3, (7, 8, 5), 1),
                                                (2, (5, 2, 11), 1), (2, (11, 6, 5), 3),
                                                (3, (8, 7, 4), 4)])
def test_sharded_matches_single_process(tmp_path, oracle, world, shape, chunks):
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(world, _free_port(), shape, str(result), chunks), nprocs=world,
             join=True)
    assert result.read_text() == "ok"


def test_world_size_one_needs_no_process_group(oracle):
    from voxelized_geometry_tools_b200.sharded import ShardedSignedDistanceField
    rng = np.random.default_rng(3)
    occupancy = (rng.random((6, 5, 4)) < 0.3).astype(np.float32)
    plan = ShardedSignedDistanceField(occupancy.shape, rank=0, world_size=1,
                                      stages=_cpu_stages(oracle))
    sdf, min_max = plan.extract(torch.from_numpy(occupancy), 0.5)
    want, (lo, hi) = oracle.sdf(occupancy, 0.5)
    np.testing.assert_array_equal(sdf.numpy(), want)
    assert (float(min_max[0]), float(min_max[1])) == (lo, hi)


# --------------------------------------------------------------------------------------------------
# Sharded voxelizer: rays split over the ranks, counters summed into x-slabs, slab-local filter
# --------------------------------------------------------------------------------------------------
def _cpu_voxelizer_stages(oracle):
    from voxelized_geometry_tools_b200.sharded import VoxelizerStages

    def raycast(points, x_gc, max_range, counts, voxel_size):
        grid = counts.numpy()
        oracle.raycast_cloud(points.numpy(), x_gc, max_range, grid.shape[:3], voxel_size,
                             counts=grid)

    def filter_slab(counts, occupancy, options):
        filtered = oracle.filter_grids(counts.numpy(), occupancy.numpy(), *options)
        occupancy.copy_(torch.from_numpy(filtered))

    return VoxelizerStages(raycast, filter_slab)


def _voxelizer_worker(rank, world, port, grid_n, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        from voxelized_geometry_tools_b200 import synthetic
        from voxelized_geometry_tools_b200.grids import compose_rigid, inverse_rigid
        from voxelized_geometry_tools_b200.sharded import (ShardedPointCloudVoxelizer,
                                                           ShardedSignedDistanceField)
        scene = synthetic.depth_camera_scene(grid_n, 5.12 / grid_n, 48, 36, max_range=5.0)
        x_gw = inverse_rigid(scene["origin_transform"])
        clouds = [(p, compose_rigid(x_gw, x), r) for p, x, r in scene["clouds"]]
        options = (0.9, 1, 2)
        dims = scene["static_occupancy"].shape
        voxelizer = ShardedPointCloudVoxelizer(dims, scene["voxel_size"],
                                               stages=_cpu_voxelizer_stages(oracle))
        x0, x1 = voxelizer.x_range
        static_slab = torch.from_numpy(scene["static_occupancy"][x0:x1].copy())
        slab = voxelizer.voxelize(static_slab, [(torch.from_numpy(p), x, r) for p, x, r in clouds],
                                  options, keep_counts=True)
        want, want_counts = oracle.voxelize(scene["static_occupancy"], clouds,
                                            scene["voxel_size"], *options)
        np.testing.assert_array_equal(slab.numpy(), want[x0:x1])
        np.testing.assert_array_equal(voxelizer.last_counts.numpy(), want_counts[:, x0:x1])
        # no clouds at all: everything that is not filled becomes unknown
        empty = voxelizer.voxelize(static_slab, [], options)
        want_empty, _ = oracle.voxelize(scene["static_occupancy"], [], scene["voxel_size"],
                                        *options)
        np.testing.assert_array_equal(empty.numpy(), want_empty[x0:x1])
        # ... and straight into the sharded SDF (the slab is what it starts from)
        plan = ShardedSignedDistanceField(dims, stages=_cpu_stages(oracle))
        sdf_slab, _ = plan.extract(slab, scene["voxel_size"])
        full = plan.gather_to_host(sdf_slab)
        if rank == 0:
            want_sdf, _ = oracle.sdf(want, scene["voxel_size"])
            np.testing.assert_array_equal(full.numpy(), want_sdf)
            Path(result_path).write_text("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,grid_n", [(2, 16), (3, 13)])
def test_sharded_voxelizer_matches_single_process(tmp_path, oracle, world, grid_n):
    result = tmp_path / "result.txt"
    mp.spawn(_voxelizer_worker, args=(world, _free_port(), grid_n, str(result)), nprocs=world,
             join=True)
    assert result.read_text() == "ok"
