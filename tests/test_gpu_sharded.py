"""GPU: the staged C-ABI passes behind the slab-sharded path, and (with >= 2 GPUs) the real
NCCL path launched under torchrun."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPO = Path(__file__).resolve().parents[1]


def test_staged_passes_emulating_four_ranks_on_one_gpu(shared_library, oracle):
    import torch
    from voxelized_geometry_tools_b200 import device as vdev
    from voxelized_geometry_tools_b200.sharded import split_range
    from .conftest import random_occupancy
    shape = (45, 38, 70)
    world = 4
    rng = np.random.default_rng(41)
    occupancy = random_occupancy(rng, shape, 0.1, blobs=True)
    dev = torch.device("cuda", 0)
    packed_slabs, from_send_layout = [], []
    for rank in range(world):
        x0, x1 = split_range(shape[0], world, rank)
        slab = torch.from_numpy(occupancy[x0:x1].copy()).to(dev)
        packed_slabs.append(vdev.edt_local_passes(slab))
        # the same passes writing send layout: block h = [nxl, rows_h, nz], back to back
        send = vdev.edt_local_passes(slab, send_parts=world).view(-1)
        blocks, offset = [], 0
        for peer in range(world):
            y0, y1 = split_range(shape[1], world, peer)
            count = (x1 - x0) * (y1 - y0) * shape[2]
            blocks.append(send[offset:offset + count].view(x1 - x0, y1 - y0, shape[2]))
            offset += count
        assert offset == send.numel()
        from_send_layout.append(torch.cat(blocks, dim=1))
    packed = torch.cat(packed_slabs, dim=0)           # what the all-to-all reassembles
    assert torch.equal(packed, torch.cat(from_send_layout, dim=0))
    for border in (False, True):
        pieces, extrema = [], []
        for rank in range(world):
            y0, y1 = split_range(shape[1], world, rank)
            y_slab = packed[:, y0:y1, :].contiguous()
            sdf, min_max = vdev.edt_final_pass(y_slab, y0, shape[1], 0.05, border)
            pieces.append(sdf)
            extrema.append(min_max)
        got = torch.cat(pieces, dim=1).cpu().numpy()
        want, (lo, hi) = oracle.sdf(occupancy, 0.05, add_virtual_border=border)
        np.testing.assert_array_equal(got, want)
        stacked = torch.stack(extrema).cpu().numpy()
        assert stacked[:, 0].min() == lo and stacked[:, 1].max() == hi


def test_device_tensor_entry_points(shared_library, oracle):
    import torch
    from voxelized_geometry_tools_b200 import device as vdev
    from .conftest import random_occupancy
    rng = np.random.default_rng(43)
    occupancy = random_occupancy(rng, (40, 50, 60), 0.2, blobs=True)
    dev = torch.device("cuda", 0)
    occ = torch.from_numpy(occupancy).to(dev)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):           # a non-default torch stream
        sdf, min_max = vdev.signed_distance_field(occ, 0.02)
        sdf64, min_max64 = vdev.signed_distance_field_f64(occ, 0.02, add_virtual_border=True)
        mask_sdf, _ = vdev.signed_distance_field_from_mask((occ > 0.5).to(torch.uint8), 0.02)
    stream.synchronize()
    want, (lo, hi) = oracle.sdf(occupancy, 0.02)
    np.testing.assert_array_equal(sdf.cpu().numpy(), want)
    assert tuple(min_max.tolist()) == (lo, hi)
    want64, (lo64, hi64) = oracle.sdf(occupancy, 0.02, add_virtual_border=True, dtype=np.float64)
    np.testing.assert_array_equal(sdf64.cpu().numpy(), want64)
    assert tuple(min_max64.tolist()) == (lo64, hi64)
    want_mask, _ = oracle.sdf_from_mask(occupancy > 0.5, 0.02)
    np.testing.assert_array_equal(mask_sdf.cpu().numpy(), want_mask)
    with pytest.raises(Exception):
        vdev.signed_distance_field(torch.from_numpy(occupancy), 0.02)   # CPU tensor: no fallback


def test_device_voxelizer_pieces(shared_library, oracle):
    import torch
    from voxelized_geometry_tools_b200 import device as vdev, synthetic
    from voxelized_geometry_tools_b200 import PointCloudVoxelizationFilterOptions
    from . import scenes
    scene = synthetic.depth_camera_scene(64, 0.06, 120, 90, max_range=5.0)
    n = 64
    dev = torch.device("cuda", 0)
    x_gw = scenes.inverse_rigid(scene["origin_transform"])
    counts = torch.zeros((len(scene["clouds"]), n, n, n, 2), dtype=torch.int32, device=dev)
    for index, (points, x_wc, max_range) in enumerate(scene["clouds"]):
        vdev.raycast_cloud(torch.from_numpy(points).to(dev), scenes.compose(x_gw, x_wc), max_range,
                           counts[index], scene["voxel_size"])
    occupancy = torch.from_numpy(scene["static_occupancy"]).to(dev)
    vdev.filter_grids(counts, occupancy, PointCloudVoxelizationFilterOptions(0.9, 2, 2))
    want, want_counts = oracle.voxelize(
        scene["static_occupancy"], [(p, scenes.compose(x_gw, x), r) for p, x, r in scene["clouds"]],
        scene["voxel_size"], 0.9, 2, 2)
    np.testing.assert_array_equal(counts.cpu().numpy(), want_counts)
    np.testing.assert_array_equal(occupancy.cpu().numpy(), want)


def test_nccl_sharded_path_on_two_gpus(shared_library):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    command = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
               "--master-addr", "127.0.0.1", "--master-port", "29531",
               str(REPO / "tests" / "multi_gpu_check.py")]
    result = subprocess.run(command, env=env, capture_output=True, text=True, timeout=600)
    assert result.returncode == 0, result.stdout[-3000:] + result.stderr[-3000:]
    assert "MULTI_GPU_CHECK_OK" in result.stdout


def test_single_process_multi_device_entry(shared_library, oracle):
    # vgt_b200_sdf_f32_multi: one process, several devices, peer stores between them; the result
    # equals the one-device call bit for bit (and the oracle at a small size).
    import torch
    import voxelized_geometry_tools_b200 as vgt
    from voxelized_geometry_tools_b200 import synthetic
    count = torch.cuda.device_count()
    if count < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    for dims, border in (((40, 36, 44), False), ((61, 45, 130), True), ((256, 192, 160), False)):
        occupancy = synthetic.clustered_spheres_occupancy(dims)
        sizes = vgt.VoxelGridSizes.FromVoxelCounts(0.02, dims)
        grid = vgt.OccupancyMap(np.eye(4), "world", sizes, data=occupancy)
        parameters = vgt.SignedDistanceFieldGenerationParameters(add_virtual_border=border)
        single = grid.ExtractSignedDistanceFieldFloat(parameters)
        for used in {2, count}:
            multi = grid.ExtractSignedDistanceFieldFloat(parameters, devices=list(range(used)))
            np.testing.assert_array_equal(multi.GetImmutableRawData(),
                                          single.GetImmutableRawData())
            assert multi.GetMinimumMaximum() == single.GetMinimumMaximum()
        if np.prod(dims) < 2 ** 20:
            want, _ = oracle.sdf(occupancy, 0.02, add_virtual_border=border)
            np.testing.assert_array_equal(multi.GetImmutableRawData(), want)
    with pytest.raises(ValueError):
        grid.ExtractSignedDistanceFieldFloat(parameters, devices=[0, 0])
