"""CPU: pins the restated oracle against the REFERENCE'S OWN SDF generation source, compiled
unmodified over shim headers (oracle/_ref, built by `make -C oracle ref` where /root/reference
exists; the prebuilt library travels to the GPU box). Skipped only if that library is absent."""
import itertools

import numpy as np
import pytest

from oracle import reference_oracle

from .conftest import occupancy_from_golden_case, random_occupancy

pytestmark = pytest.mark.skipif(not reference_oracle.available(),
                                reason="oracle/_ref/libvgt_ref.so not built")


def test_reference_build_reproduces_its_own_goldens(sdf_goldens):
    tolerance = sdf_goldens["extrema_tolerance"]
    for case in sdf_goldens["cases"]:
        occupancy, resolution = occupancy_from_golden_case(case)
        sdf, extrema = reference_oracle.sdf(occupancy, resolution)
        if case.get("expected_min_max"):
            for got, want in zip(extrema, case["expected_min_max"]):
                want = float(want)
                assert got == want or abs(got - want) <= tolerance
        for cell in case["expected_cells"]:
            x, y, z = cell["index"]
            assert sdf[x, y, z] == np.float32(cell["value"])


@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 1, 12), (4, 8, 12), (9, 1, 7), (13, 11, 17),
                                   (3, 40, 9), (33, 20, 25)])
def test_restatement_equals_reference(oracle, shape):
    rng = np.random.default_rng(abs(hash(shape)) % 2 ** 32)
    for fill, unknown_is_filled, border in itertools.product(
            (0.0, 0.05, 0.5, 1.0), (True, False), (False, True)):
        occupancy = random_occupancy(rng, shape, fill)
        for dtype in (np.float32, np.float64):
            mine, mine_extrema = oracle.sdf(occupancy, 0.3, unknown_is_filled, border, dtype=dtype)
            theirs, their_extrema = reference_oracle.sdf(occupancy, 0.3, unknown_is_filled, border,
                                                         dtype=dtype)
            np.testing.assert_array_equal(mine, theirs)
            assert mine_extrema == their_extrema


@pytest.mark.skipif(not reference_oracle.maps_available(),
                    reason="oracle/_ref/libvgt_ref_maps.so not built")
@pytest.mark.parametrize("shape", [(1, 1, 1), (4, 8, 12), (9, 1, 7), (13, 11, 17), (3, 40, 9)])
def test_restatement_equals_the_references_own_occupancy_map_member(oracle, shape):
    # OccupancyMap::ExtractSignedDistanceField<T> itself (occupancy_map.hpp:174-216): the public
    # entry, with the reference's own predicate (:181-205) on its own OccupancyCell grid.
    rng = np.random.default_rng(abs(hash(shape)) % 2 ** 32 + 5)
    for fill, unknown_is_filled, border in itertools.product(
            (0.0, 0.1, 0.6, 1.0), (True, False), (False, True)):
        occupancy = random_occupancy(rng, shape, fill, unknown=0.2)
        for dtype in (np.float32, np.float64):
            mine, mine_extrema = oracle.sdf(occupancy, 0.3, unknown_is_filled, border, dtype=dtype)
            theirs, their_extrema = reference_oracle.occupancy_map_sdf(
                occupancy, 0.3, unknown_is_filled, border, dtype=dtype)
            np.testing.assert_array_equal(mine, theirs)
            assert mine_extrema == their_extrema


def test_restatement_equals_reference_medium(oracle):
    rng = np.random.default_rng(101)
    occupancy = random_occupancy(rng, (64, 72, 80), 0.1, blobs=True)
    for threads in (1, 0):
        mine, _ = oracle.sdf(occupancy, 0.02, threads=threads)
        theirs, _ = reference_oracle.sdf(occupancy, 0.02, threads=threads)
        np.testing.assert_array_equal(mine, theirs)


def test_transform_in_place_on_arbitrary_samples(oracle):
    # ComputeDistanceFieldTransformInPlace on non-binary sampled functions (integers + inf).
    rng = np.random.default_rng(7)
    field = rng.integers(0, 400, size=(11, 14, 9)).astype(np.float64)
    field[rng.random(field.shape) < 0.3] = np.inf
    mine = oracle.transform_inplace(field.copy())
    theirs = reference_oracle.transform_inplace(field.copy())
    np.testing.assert_array_equal(mine, theirs)


# ------------------------------------------------------------------------------------------------
# Voxelizer: the restated oracle against the REFERENCE'S OWN cpu_pointcloud_voxelization.cpp
# (compiled unmodified into oracle/_ref). Raw seen-free / seen-filled counts, bit for bit.
# ------------------------------------------------------------------------------------------------
needs_reference_voxelizer = pytest.mark.skipif(
    not reference_oracle.voxelizer_available(),
    reason="oracle/_ref/libvgt_ref.so was built without the voxelizer")


@needs_reference_voxelizer
def test_reference_voxelizer_reproduces_its_own_test_scene():
    # test/pointcloud_voxelization_test.cpp:84-158 through VoxelizePointClouds itself (posed grid).
    from . import scenes
    scene = scenes.reference_voxelization_scene()
    empty = reference_oracle.voxelize_posed(scene["static"], scene["x_wg"], [],
                                            scene["voxel_size"], *scene["filter"])
    scenes.check_empty_voxelization(empty)
    filtered = reference_oracle.voxelize_posed(scene["static"], scene["x_wg"], scene["clouds"],
                                               scene["voxel_size"], *scene["filter"])
    scenes.check_voxelization(filtered)


@needs_reference_voxelizer
def test_voxelizer_counts_equal_reference_on_its_test_scene(oracle):
    from voxelized_geometry_tools_b200.grids import compose_rigid, inverse_rigid
    from . import scenes
    scene = scenes.reference_voxelization_scene()
    x_gw = inverse_rigid(scene["x_wg"])
    prepared = [(p, compose_rigid(x_gw, x), r) for p, x, r in scene["clouds"]]
    for threads in (1, 0):
        mine, my_counts = oracle.voxelize(scene["static"], prepared, scene["voxel_size"],
                                          *scene["filter"], threads=threads)
        theirs, their_counts = reference_oracle.voxelize(
            scene["static"], prepared, scene["voxel_size"], *scene["filter"], threads=threads)
        np.testing.assert_array_equal(my_counts, their_counts)
        np.testing.assert_array_equal(mine, theirs)
    # and the posed interface call (the reference composes X_GC itself) gives the same map
    posed = reference_oracle.voxelize_posed(scene["static"], scene["x_wg"], scene["clouds"],
                                            scene["voxel_size"], *scene["filter"])
    np.testing.assert_array_equal(posed, theirs)


@needs_reference_voxelizer
def test_single_rays_equal_reference_on_the_seeded_pairs(oracle):
    # the 1000 mt19937_64(42) origin / point pairs of test/voxel_raycasting_test.cpp:61-100,
    # unclipped and clipped (max ranges that cut most rays short)
    from . import scenes
    g, pairs = scenes.random_ray_pairs()
    dims = tuple(g["voxel_counts"])
    for max_range in (g["max_range"], 3.0, 0.7):
        for origin, point in pairs:
            mine = oracle.raycast_single(origin, point, max_range, dims, g["resolution"])
            theirs = reference_oracle.raycast_single(origin, point, max_range, dims,
                                                     g["resolution"])
            assert np.array_equal(mine, theirs), (origin, point, max_range)


@needs_reference_voxelizer
def test_voxelizer_counts_equal_reference_on_camera_scenes(oracle):
    # BASELINE config 3 at reduced size: four posed pinhole cameras, NaN pixels, clipped rays,
    # rays that leave the grid; two filter settings. Then rays from outside the grid.
    from voxelized_geometry_tools_b200 import synthetic
    from voxelized_geometry_tools_b200.grids import compose_rigid, inverse_rigid
    for grid_n, voxel, width, height, max_range in ((64, 0.08, 160, 120, 6.0),
                                                    (96, 0.04, 120, 90, 2.5),
                                                    (40, 0.1, 64, 48, float("inf"))):
        scene = synthetic.depth_camera_scene(grid_n, voxel, width, height, max_range=max_range)
        x_gw = inverse_rigid(scene["origin_transform"])
        prepared = [(p, compose_rigid(x_gw, x), r) for p, x, r in scene["clouds"]]
        for options in ((1.0, 1, 1), (0.9, 2, 2)):
            mine, my_counts = oracle.voxelize(scene["static_occupancy"], prepared, voxel, *options)
            theirs, their_counts = reference_oracle.voxelize(scene["static_occupancy"], prepared,
                                                             voxel, *options)
            np.testing.assert_array_equal(my_counts, their_counts)
            np.testing.assert_array_equal(mine, theirs)
        assert int(their_counts.sum()) > 100000
        posed = reference_oracle.voxelize_posed(scene["static_occupancy"],
                                                scene["origin_transform"], scene["clouds"], voxel,
                                                0.9, 2, 2)
        np.testing.assert_array_equal(posed, theirs)


@needs_reference_voxelizer
def test_voxelizer_counts_equal_reference_on_random_posed_clouds(oracle):
    # random rigid poses (origins inside and far outside the grid), random points, finite and
    # infinite max range, some non-finite points, zero-length rays
    rng = np.random.default_rng(2024)
    dims, voxel = (24, 20, 28), 0.05
    static = np.zeros(dims, dtype=np.float32)
    for trial in range(30):
        angles = rng.uniform(-np.pi, np.pi, 3)
        cx, cy, cz = np.cos(angles)
        sx, sy, sz = np.sin(angles)
        rotation = np.array([[cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx],
                             [sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx],
                             [-sy, cy * sx, cy * cx]])
        x_gc = np.eye(4)
        x_gc[:3, :3] = rotation
        x_gc[:3, 3] = rng.uniform(-1.5, 2.5, 3) if trial % 3 else rng.uniform(0.1, 0.9, 3)
        points = rng.uniform(-3.0, 3.0, (4000, 3))
        points[rng.random(4000) < 0.01] = np.nan
        points[rng.random(4000) < 0.01, 2] = np.inf
        points[:5] = 0.0                                 # zero-length rays
        max_range = float(rng.choice([np.inf, 4.0, 1.0]))
        mine, my_counts = oracle.voxelize(static, [(points, x_gc, max_range)], voxel, 0.9, 1, 1)
        theirs, their_counts = reference_oracle.voxelize(static, [(points, x_gc, max_range)],
                                                         voxel, 0.9, 1, 1)
        np.testing.assert_array_equal(my_counts, their_counts, err_msg=f"trial {trial}")
        np.testing.assert_array_equal(mine, theirs)


@needs_reference_voxelizer
def test_front_end_pose_composition_matches_the_stand_in():
    # X_GC = X_GW * X_WC is composed by the front ends; its last bits decide boundary voxels.
    from voxelized_geometry_tools_b200.grids import compose_rigid, inverse_rigid
    rng = np.random.default_rng(5)
    for _ in range(200):
        def pose():
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            w, x, y, z = q
            m = np.eye(4)
            m[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
            m[:3, 3] = rng.uniform(-5, 5, 3)
            return m
        a, b = pose(), pose()
        assert np.array_equal(compose_rigid(a, b), reference_oracle.isometry_product(a, b))
        assert np.array_equal(inverse_rigid(a), reference_oracle.isometry_inverse(a))
