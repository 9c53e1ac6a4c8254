"""CPU: oracle/mesh_rasterizer_oracle.cpp (the restatement the CUDA rasterizer is checked against)
against the reference's own mesh_rasterizer.cpp compiled unmodified into oracle/_ref, bit for bit,
and against the expectations of the reference's own test (test/mesh_rasterization_test.cpp)."""
import numpy as np
import pytest

from oracle import oracle as restated, reference_oracle
from tests import meshes

needs_reference = pytest.mark.skipif(
    not reference_oracle.available(),
    reason="oracle/_ref was not built (needs /root/reference at build time)")


def _check_reference_test_triangle(occupancy):
    """test/mesh_rasterization_test.cpp:37-65."""
    nx, ny, nz = occupancy.shape
    assert np.all(occupancy[:, :, 0] == 0.0)
    for x in range(nx):
        for y in range(ny):
            want = 0.0 if (x == 0 or y == 0 or y >= ny - x) else 1.0
            assert occupancy[x, y, 1] == want, (x, y)


def test_restatement_reproduces_the_reference_test_triangle():
    vertices = [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]
    occupancy, origin, code = restated.rasterize_mesh_into_occupancy_map(vertices, [[0, 1, 2]],
                                                                         0.125)
    assert code == restated.RASTERIZE_OK
    assert occupancy.shape == (10, 10, 2)
    np.testing.assert_array_equal(origin[:3, 3], [-0.125, -0.125, -0.125])
    _check_reference_test_triangle(occupancy)


@needs_reference
def test_reference_build_reproduces_its_own_test_triangle():
    vertices = [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]
    for threads in (1, 0):
        occupancy, origin, code = reference_oracle.rasterize_mesh_into_occupancy_map(
            vertices, [[0, 1, 2]], 0.125, threads)
        assert code == 0
        _check_reference_test_triangle(occupancy)


@needs_reference
@pytest.mark.parametrize("name", ["icosphere", "random_soup", "slivers", "degenerate", "box"])
@pytest.mark.parametrize("resolution", [0.05, 0.0625])
def test_into_occupancy_map_equals_reference(name, resolution):
    vertices, triangles = meshes.make(name)
    got, got_origin, got_code = restated.rasterize_mesh_into_occupancy_map(vertices, triangles,
                                                                          resolution)
    want, want_origin, want_code = reference_oracle.rasterize_mesh_into_occupancy_map(
        vertices, triangles, resolution)
    assert got_code == want_code == 0
    np.testing.assert_array_equal(got_origin, want_origin)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)
    assert 0 < got.sum() < got.size


@needs_reference
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_into_a_posed_map_equals_reference(seed):
    """RasterizeMesh into an existing map with a rotated origin, mesh partly outside."""
    rng = np.random.default_rng(seed)
    vertices, triangles = meshes.make("random_soup", seed)
    dims, resolution = (24, 20, 28), 0.05
    origin = meshes.pose_centred_on_origin(rng, dims, resolution)
    base = (rng.random(dims) < 0.05).astype(np.float32) * 0.5
    got, want = base.copy(), base.copy()
    got_code = restated.rasterize_mesh(vertices, triangles, got, resolution, origin, False)
    want_code = reference_oracle.rasterize_mesh(vertices, triangles, want, resolution, origin,
                                                False, threads=0)
    assert got_code == want_code == 0
    np.testing.assert_array_equal(got, want)
    assert (got == 1.0).sum() > 50
    # enforce: the mesh leaves this map, the reference throws (serial loop: thrown exceptions
    # cannot leave an OpenMP region)
    assert restated.rasterize_mesh(vertices, triangles, base.copy(), resolution, origin,
                                   True) == restated.RASTERIZE_NOT_CONTAINED
    assert reference_oracle.rasterize_mesh(vertices, triangles, base.copy(), resolution, origin,
                                           True, threads=1) == 2


@needs_reference
def test_bad_vertex_index_is_an_error_in_both():
    vertices, triangles = meshes.make("box")
    triangles = triangles.copy()
    triangles[3, 1] = len(vertices)
    occupancy = np.zeros((8, 8, 8), dtype=np.float32)
    assert restated.rasterize_mesh(vertices, triangles, occupancy, 0.2) \
        == restated.RASTERIZE_BAD_INDEX
    assert reference_oracle.rasterize_mesh(vertices, triangles, occupancy, 0.2, threads=1) == 3
