"""CPU: checks the voxelizer oracle against everything the reference's tests pin for it."""
import numpy as np
import pytest

from . import scenes


def _oracle_voxelize(oracle, scene, clouds):
    x_gw = scenes.inverse_rigid(scene["x_wg"])
    prepared = [(points, scenes.compose(x_gw, x_wc), max_range) for points, x_wc, max_range in clouds]
    return oracle.voxelize(scene["static"], prepared, scene["voxel_size"], *scene["filter"])


def test_reference_scene_planes(oracle):
    scene = scenes.reference_voxelization_scene()
    assert scene["static"].shape == (8, 8, 8)
    assert scene["clouds"][0][0].shape == (129 * 129, 3)
    empty, _ = _oracle_voxelize(oracle, scene, [])
    scenes.check_empty_voxelization(empty)
    filtered, counts = _oracle_voxelize(oracle, scene, scene["clouds"])
    scenes.check_voxelization(filtered)
    assert counts.shape == (3, 8, 8, 8, 2)
    assert counts[2].sum() == 0  # the empty cloud


@pytest.mark.parametrize("threads", [1, 4])
def test_threads_do_not_change_counts(oracle, threads):
    scene = scenes.reference_voxelization_scene()
    x_gw = scenes.inverse_rigid(scene["x_wg"])
    points, x_wc, max_range = scene["clouds"][0]
    base = oracle.raycast_cloud(points, scenes.compose(x_gw, x_wc), max_range, (8, 8, 8), 0.25, threads=1)
    other = oracle.raycast_cloud(points, scenes.compose(x_gw, x_wc), max_range, (8, 8, 8), 0.25, threads=threads)
    np.testing.assert_array_equal(base, other)


def test_each_voxel_at_most_once_per_ray(oracle):
    # test/voxel_raycasting_test.cpp:61-82.
    g, pairs = scenes.random_ray_pairs()
    dims = tuple(g["voxel_counts"])
    touched = 0
    for origin, point in pairs:
        counts = oracle.raycast_single(origin, point, g["max_range"], dims, g["resolution"])
        assert counts.min() >= 0 and counts.max() <= 1
        assert not np.any((counts[..., 0] > 0) & (counts[..., 1] > 0))
        touched += int(counts.sum())
    assert touched > 10000  # the rays do cross the grid


def test_non_finite_single_ray_is_rejected(oracle):
    # cpu_pcv.cpp:91-99 throws invalid_argument.
    with pytest.raises(ValueError):
        oracle.raycast_single([np.nan, 0, 0], [1, 1, 1], 10.0, (4, 4, 4), 0.5)


def test_non_finite_points_are_skipped(oracle):
    points = np.array([[np.nan, 0.0, 1.0], [0.1, 0.1, np.inf], [0.3, 0.3, 0.3]])
    counts = oracle.raycast_cloud(points, np.eye(4), np.inf, (4, 4, 4), 0.25)
    assert counts[..., 1].sum() == 1


def test_clipped_ray_marks_final_voxel_free(oracle):
    # cpu_pcv.cpp:371-375: a clipped ray ends in a seen-free voxel.
    x_gc = scenes.translation(0.125, 0.125, 0.125)
    points = np.array([[3.0, 0.0, 0.0]])
    counts = oracle.raycast_cloud(points, x_gc, 1.0, (8, 8, 8), 0.25)
    assert counts[..., 1].sum() == 0
    np.testing.assert_array_equal(np.flatnonzero(counts[:, 0, 0, 0]), np.arange(0, 5))


def test_filter_rule(oracle):
    # pcv_if.hpp:55-86 + cpu_pcv.cpp:453-488 on hand-made counts.
    counts = np.zeros((2, 1, 1, 6, 2), dtype=np.int32)
    counts[0, 0, 0, 0] = (3, 0)      # free only                -> 0.0
    counts[0, 0, 0, 1] = (0, 2)      # filled only              -> 1.0
    counts[0, 0, 0, 2] = (9, 1)      # 90 % free, thr 0.9       -> free
    counts[0, 0, 0, 3] = (8, 2)      # 80 % free                -> filled
    counts[0, 0, 0, 4] = (0, 0)      # unseen                   -> 0.5
    counts[1, 0, 0, 5] = (1, 0)      # seen free by camera 2 only
    static = np.zeros((1, 1, 6), dtype=np.float32)
    out = oracle.filter_grids(counts, static, 0.9, 1, 1)
    np.testing.assert_array_equal(out.reshape(-1), [0.0, 1.0, 0.0, 1.0, 0.5, 0.0])
    out2 = oracle.filter_grids(counts, static, 0.9, 2, 2)   # outlier thr 2, two cameras needed
    np.testing.assert_array_equal(out2.reshape(-1), [0.5, 1.0, 0.5, 1.0, 0.5, 0.5])
    static_filled = np.full((1, 1, 6), 0.75, dtype=np.float32)
    np.testing.assert_array_equal(oracle.filter_grids(counts, static_filled, 0.9, 1, 1),
                                  static_filled)
    for bad in ((0.0, 1, 1), (1.5, 1, 1), (0.5, 0, 1), (0.5, 1, 0)):
        with pytest.raises(ValueError):
            oracle.filter_grids(counts, static, *bad)


def test_origin_outside_grid_enters_through_slab(oracle):
    # cpu_pcv.cpp:229-290: origin outside, ray crosses the grid along +x through row (y=1, z=1).
    x_gc = scenes.translation(-1.0, 0.375, 0.375)
    counts = oracle.raycast_cloud(np.array([[1.6, 0.0, 0.0]]), x_gc, np.inf, (4, 4, 4), 0.25)
    np.testing.assert_array_equal(counts[:, 1, 1, 0], [1, 1, 0, 0])
    np.testing.assert_array_equal(counts[:, 1, 1, 1], [0, 0, 1, 0])
    assert counts.sum() == 3
    # A ray that misses the grid entirely leaves nothing behind.
    miss = oracle.raycast_cloud(np.array([[0.0, 5.0, 0.0]]), x_gc, np.inf, (4, 4, 4), 0.25)
    assert miss.sum() == 0
