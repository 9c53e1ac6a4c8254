"""GPU: the window kernel (csrc/edt_envelope_window.cuh) on each of its three routes - rows that
certify inside the register window, rows that take the extended search, tiles handed to the stack
kernel - against the oracle, bit for bit. The routes are forced with the library's A/B switches
(read at every call): VGT_B200_WINDOW_BUDGET (extended-search steps a warp may spend per 100 rows of
its tile; 0 sends every tile with an uncertain row to the stack kernel, a huge value
keeps everything in the window kernel), VGT_B200_WINDOW_PILOT=0 (no pilot launch: the window
kernel takes every tile whatever the map looks like), VGT_B200_WINDOW_STAGE=0 / 1 (rows prefetched
into registers / staged through shared memory with cp.async, for both passes) and VGT_B200_ENVELOPE=lean (no
window kernel at all)."""
import os

import numpy as np
import pytest

from voxelized_geometry_tools_b200 import synthetic

from .conftest import random_occupancy
from .test_gpu_sdf import assert_matches_oracle

pytestmark = pytest.mark.gpu

ROUTES = [{"VGT_B200_WINDOW_BUDGET": "0"}, {"VGT_B200_WINDOW_BUDGET": "100000"},
          {"VGT_B200_WINDOW_BUDGET": "25"}, {"VGT_B200_ENVELOPE": "lean"},
          {"VGT_B200_WINDOW_PILOT": "0"}, {"VGT_B200_WINDOW_STAGE": "1"},
          {"VGT_B200_WINDOW_STAGE": "2"}, {"VGT_B200_WINDOW_STAGE": "2", "VGT_B200_WINDOW_PILOT": "0"},
          {"VGT_B200_WINDOW_STAGE": "0", "VGT_B200_WINDOW_PILOT": "0"},
          {"VGT_B200_WINDOW_RADIUS_Y": "8"}, {"VGT_B200_WINDOW_RADIUS_Y": "10"}, {}]


@pytest.fixture(params=ROUTES, ids=lambda route: ",".join(f"{k}={v}" for k, v in route.items())
                or "default")
def route(request):
    from voxelized_geometry_tools_b200 import _capi
    saved = {key: os.environ.get(key) for key in ("VGT_B200_WINDOW_BUDGET", "VGT_B200_ENVELOPE",
                                                  "VGT_B200_WINDOW_PILOT", "VGT_B200_WINDOW_STAGE",
                                                  "VGT_B200_WINDOW_RADIUS_Y")}
    for key in saved:
        os.environ.pop(key, None)
    os.environ.update(request.param)
    _capi.library().vgt_b200_reload_tuning()    # the switches are read once per process
    yield request.param
    for key, value in saved.items():
        os.environ.pop(key, None)
        if value is not None:
            os.environ[key] = value
    _capi.library().vgt_b200_reload_tuning()


def test_cluttered_grid(shared_library, oracle, route):
    # distances of a few voxels nearly everywhere: the certified route
    occupancy = synthetic.clustered_spheres_occupancy((96, 80, 72))
    assert_matches_oracle(oracle, occupancy, 0.02)
    assert_matches_oracle(oracle, occupancy, 0.02, add_virtual_border=True)
    assert_matches_oracle(oracle, occupancy, 0.02, dtype=np.float64)


def test_sparse_and_empty_grids(shared_library, oracle, route):
    # distances far beyond the window: extended search or the stack kernel; an empty and a full
    # grid (no opposite class anywhere: every value is infinite)
    rng = np.random.default_rng(5)
    occupancy = np.zeros((70, 45, 50), dtype=np.float32)
    occupancy[rng.integers(0, 70, 6), rng.integers(0, 45, 6), rng.integers(0, 50, 6)] = 1.0
    assert_matches_oracle(oracle, occupancy, 0.05)
    assert_matches_oracle(oracle, 1.0 - occupancy, 0.05, add_virtual_border=True)
    assert_matches_oracle(oracle, np.zeros((13, 30, 17), dtype=np.float32), 0.05)
    assert_matches_oracle(oracle, np.ones((13, 30, 17), dtype=np.float32), 0.05,
                          add_virtual_border=True)


@pytest.mark.parametrize("shape", [(1, 1, 5), (2, 3, 4), (7, 1, 33), (1, 13, 40), (12, 24, 36),
                                   (13, 25, 37), (11, 23, 35), (8, 16, 33), (9, 17, 31),
                                   (40, 100, 31), (100, 37, 65)])
def test_line_lengths_around_the_chunk_size(shared_library, oracle, route, shape):
    # chunk sizes are 12 (y) and 8 (x): lengths below, at and just past multiples of them,
    # column counts that are not multiples of 32 (shadow lanes)
    rng = np.random.default_rng(sum(shape))
    for fill in (0.03, 0.4):
        occupancy = random_occupancy(rng, shape, fill)
        assert_matches_oracle(oracle, occupancy, 0.1)
        assert_matches_oracle(oracle, occupancy, 0.1, add_virtual_border=True, dtype=np.float64)


def test_box_scene_and_blobs(shared_library, oracle, route):
    assert_matches_oracle(oracle, synthetic.box_scene(64), 0.02)
    rng = np.random.default_rng(11)
    occupancy = random_occupancy(rng, (60, 130, 90), 0.08, blobs=True)
    assert_matches_oracle(oracle, occupancy, 0.02)


@pytest.mark.parametrize("shape,world", [((20, 38, 70), 4), ((9, 130, 40), 3), ((6, 61, 33), 8)])
def test_send_layout_output(shared_library, route, shape, world):
    # the slab-local y pass writing the all-to-all's send layout (multi-GPU path) must hold the
    # packed result of the plain pass, part by part
    import torch
    from voxelized_geometry_tools_b200 import device as vdev
    from voxelized_geometry_tools_b200.sharded import split_range
    rng = np.random.default_rng(world)
    for fill in (0.02, 0.3):
        occupancy = random_occupancy(rng, shape, fill, blobs=True)
        slab = torch.from_numpy(occupancy).to(torch.device("cuda", 0))
        packed = vdev.edt_local_passes(slab)
        send = vdev.edt_local_passes(slab, send_parts=world).view(-1)
        blocks, offset = [], 0
        for peer in range(world):
            y0, y1 = split_range(shape[1], world, peer)
            count = shape[0] * (y1 - y0) * shape[2]
            blocks.append(send[offset:offset + count].view(shape[0], y1 - y0, shape[2]))
            offset += count
        assert offset == send.numel()
        assert torch.equal(packed, torch.cat(blocks, dim=1))


def test_grids_with_enough_tiles_for_the_pilot(shared_library, oracle, route):
    # >= 64 groups of 4 tiles in both passes: the pilot launch decides between "cluttered" (the
    # rest of the window launch runs) and "deep" (the stack kernel takes every tile)
    shape = (40, 100, 256)
    cluttered = synthetic.clustered_spheres_occupancy(shape)
    assert_matches_oracle(oracle, cluttered, 0.02)
    rng = np.random.default_rng(3)
    deep = np.zeros(shape, dtype=np.float32)
    deep[rng.integers(0, 40, 5), rng.integers(0, 100, 5), rng.integers(0, 256, 5)] = 1.0
    deep[:, 60:, 200:] = 1.0
    assert_matches_oracle(oracle, deep, 0.02, add_virtual_border=True)
    mixed = cluttered.copy()
    mixed[:, :50, :] = 0.0
    assert_matches_oracle(oracle, mixed, 0.02)
