"""GPU: the CUDA mesh rasterizer through the C-ABI / the Python mirror of the reference's
mesh_rasterizer interface against oracle/mesh_rasterizer_oracle.cpp (itself pinned bit for bit to
the reference's own mesh_rasterizer.cpp, tests/test_mesh_rasterizer_vs_reference.py): the same
cells, bit for bit, and the same errors."""
import ctypes

import numpy as np
import pytest

from tests import meshes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rasterizer(shared_library):
    from voxelized_geometry_tools_b200 import mesh_rasterizer
    return mesh_rasterizer


def test_the_reference_test_triangle(rasterizer):
    """test/mesh_rasterization_test.cpp:20-66 (and :68-114 for the component map)."""
    vertices = [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]
    plain = rasterizer.RasterizeMeshIntoOccupancyMap(vertices, [[0, 1, 2]], 0.125)
    component = rasterizer.RasterizeMeshIntoOccupancyComponentMap(vertices, [[0, 1, 2]], 0.125)
    for occupancy in (plain.GetImmutableRawData(),
                      component.GetImmutableRawData()["occupancy"]):
        nx, ny, nz = occupancy.shape
        assert (nx, ny, nz) == (10, 10, 2)
        assert np.all(occupancy[:, :, 0] == 0.0)
        for x in range(nx):
            for y in range(ny):
                want = 0.0 if (x == 0 or y == 0 or y >= ny - x) else 1.0
                assert occupancy[x, y, 1] == want, (x, y)
    assert np.all(component.GetImmutableRawData()["component"] == 0)


@pytest.mark.parametrize("name", ["icosphere", "random_soup", "slivers", "degenerate", "box"])
@pytest.mark.parametrize("resolution", [0.05, 0.0625, 0.013])
def test_into_occupancy_map_equals_oracle(rasterizer, name, resolution):
    from oracle import oracle
    vertices, triangles = meshes.make(name)
    got = rasterizer.RasterizeMeshIntoOccupancyMap(vertices, triangles, resolution)
    want, want_origin, code = oracle.rasterize_mesh_into_occupancy_map(vertices, triangles,
                                                                       resolution)
    assert code == oracle.RASTERIZE_OK
    np.testing.assert_array_equal(got.OriginTransform(), want_origin)
    assert got.GetImmutableRawData().shape == want.shape
    np.testing.assert_array_equal(got.GetImmutableRawData(), want)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_into_a_posed_map_equals_oracle_and_keeps_other_cells(rasterizer, seed):
    from oracle import oracle
    from voxelized_geometry_tools_b200.grids import (OccupancyComponentMap, OccupancyMap,
                                                     VoxelGridSizes)
    rng = np.random.default_rng(seed)
    vertices, triangles = meshes.make("random_soup", seed)
    dims, resolution = (24, 20, 28), 0.05
    origin = meshes.pose_centred_on_origin(rng, dims, resolution,
                                           max_angle=0.15 if seed else 0.0)
    base = (rng.random(dims) < 0.05).astype(np.float32) * 0.5
    want = base.copy()
    assert oracle.rasterize_mesh(vertices, triangles, want, resolution, origin, False) == 0
    assert (want == 1.0).sum() > 50
    sizes = VoxelGridSizes.FromVoxelCounts(resolution, dims)
    grid = OccupancyMap(origin, "world", sizes, data=base.copy())
    rasterizer.RasterizeMesh(vertices, triangles, grid, False)
    np.testing.assert_array_equal(grid.GetImmutableRawData(), want)
    # component map: 8-byte cells, the component word is left alone
    cells = OccupancyComponentMap(origin, "world", sizes)
    cells.GetMutableRawData()["occupancy"] = base
    cells.GetMutableRawData()["component"] = rng.integers(0, 2 ** 32, dims, dtype=np.uint32)
    components = cells.GetImmutableRawData()["component"].copy()
    rasterizer.RasterizeMesh(vertices, triangles, cells, False)
    np.testing.assert_array_equal(cells.GetImmutableRawData()["occupancy"], want)
    np.testing.assert_array_equal(cells.GetImmutableRawData()["component"], components)
    # the mesh leaves this map: with enforcement the reference throws std::runtime_error
    with pytest.raises(RuntimeError, match="not contained"):
        rasterizer.RasterizeMesh(vertices, triangles, OccupancyMap(origin, "world", sizes), True)
    # triangle by triangle == the whole mesh; rasterizing twice changes nothing
    one_by_one = OccupancyMap(origin, "world", sizes, data=base.copy())
    for index in range(0, len(triangles), 7):
        rasterizer.RasterizeTriangle(vertices, triangles, index, one_by_one, False)
    subset = base.copy()
    oracle.rasterize_mesh(vertices, triangles[::7], subset, resolution, origin, False)
    np.testing.assert_array_equal(one_by_one.GetImmutableRawData(), subset)
    rasterizer.RasterizeMesh(vertices, triangles, grid, False)
    np.testing.assert_array_equal(grid.GetImmutableRawData(), want)


def test_errors_follow_the_reference(rasterizer):
    from voxelized_geometry_tools_b200.grids import OccupancyMap, VoxelGridSizes
    vertices, triangles = meshes.make("box")
    grid = OccupancyMap(np.eye(4), "world", VoxelGridSizes.FromVoxelCounts(0.2, (8, 8, 8)))
    bad = triangles.copy()
    bad[3, 1] = len(vertices)
    with pytest.raises(IndexError):
        rasterizer.RasterizeMesh(vertices, bad, grid, False)
    bad[3, 1] = -1
    with pytest.raises(IndexError):
        rasterizer.RasterizeMesh(vertices, bad, grid, False)
    with pytest.raises(IndexError):
        rasterizer.RasterizeTriangle(vertices, triangles, len(triangles), grid, False)
    with pytest.raises(ValueError):
        rasterizer.RasterizeMeshIntoOccupancyMap(vertices, triangles, 0.0)
    with pytest.raises(ValueError):
        rasterizer.RasterizeMesh(vertices[:, :2], triangles, grid, False)
    # an empty triangle list is fine and changes nothing (a fresh map: after an error the map
    # holds whatever was rasterized before it, in the reference as here)
    fresh = OccupancyMap(np.eye(4), "world", VoxelGridSizes.FromVoxelCounts(0.2, (8, 8, 8)))
    rasterizer.RasterizeMesh(vertices, np.zeros((0, 3), dtype=np.int32), fresh, True)
    assert not fresh.GetImmutableRawData().any()


def test_device_entry_on_a_large_mesh_then_sdf(shared_library):
    """An icosphere of 20480 triangles into a 200^3 device-resident map through the device entry,
    equal to the oracle, then straight into the SDF path (the rasterizer is an occupancy
    producer in front of it)."""
    import torch
    from oracle import oracle
    from voxelized_geometry_tools_b200 import _capi, device as vdev
    vertices, triangles = meshes.icosphere(subdivisions=5, radius=0.8, centre=(1.0, 1.0, 1.0))
    assert len(triangles) == 20480
    dims, resolution = (200, 200, 200), 0.01
    want = np.zeros(dims, dtype=np.float32)
    assert oracle.rasterize_mesh(vertices, triangles, want, resolution, np.eye(4), True) == 0
    dev = torch.device("cuda", 0)
    d_vertices = torch.from_numpy(vertices).to(dev)
    d_triangles = torch.from_numpy(triangles).to(dev)
    occupancy = torch.zeros(dims, dtype=torch.float32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    identity = np.ascontiguousarray(np.eye(4).T).reshape(16)
    pointer = identity.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    lib = _capi.library()
    before = lib.vgt_b200_kernel_launch_count()
    code = lib.vgt_b200_rasterize_mesh_dev(
        d_vertices.data_ptr(), len(vertices), d_triangles.data_ptr(), len(triangles),
        occupancy.data_ptr(), 4, *dims, resolution, pointer, pointer, 1, 0, flags.data_ptr(),
        torch.cuda.current_stream(dev).cuda_stream)
    _capi.check(code)
    assert lib.vgt_b200_kernel_launch_count() == before + 1
    assert lib.vgt_b200_rasterize_status(int(flags.item())) == 0
    np.testing.assert_array_equal(occupancy.cpu().numpy(), want)
    sdf, _ = vdev.signed_distance_field(occupancy, resolution)
    reference_sdf, _ = oracle.sdf(want, resolution)
    np.testing.assert_array_equal(sdf.cpu().numpy(), reference_sdf)
    # a watertight shell: the centre of the sphere is free space well away from the surface
    assert sdf[100, 100, 100].item() > 0.7
