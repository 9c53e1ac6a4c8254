"""GPU: the CUDA voxelizer through the C-ABI against the oracle: counts bit-exact, occupancy exact."""
import numpy as np
import pytest

import voxelized_geometry_tools_b200 as vgt
from voxelized_geometry_tools_b200 import _capi, synthetic

from . import scenes

pytestmark = pytest.mark.gpu


def build(scene, clouds):
    sizes = vgt.VoxelGridSizes.FromVoxelCounts(scene["voxel_size"], scene["static"].shape)
    static = vgt.OccupancyMap(scene["x_wg"], "world", sizes, data=scene["static"].copy())
    wrappers = [vgt.VectorPointCloudWrapper(points, x_wc, max_range)
                for points, x_wc, max_range in clouds]
    return static, wrappers


def oracle_voxelize(oracle, scene, clouds, filter_options):
    x_gw = scenes.inverse_rigid(scene["x_wg"])
    prepared = [(points, scenes.compose(x_gw, x_wc), max_range) for points, x_wc, max_range in clouds]
    return oracle.voxelize(scene["static"], prepared, scene["voxel_size"], *filter_options)


def test_reference_scene(shared_library, oracle):
    # test/pointcloud_voxelization_test.cpp:269-295 for this backend.
    scene = scenes.reference_voxelization_scene()
    options = vgt.PointCloudVoxelizationFilterOptions(*scene["filter"])
    for backend in vgt.GetAvailableBackends():
        messages = []
        voxelizer = vgt.MakePointCloudVoxelizer(backend["options"], messages.append)
        static, wrappers = build(scene, scene["clouds"])
        runtimes = []
        empty = voxelizer.VoxelizePointClouds(static, options, [], runtimes.append)
        scenes.check_empty_voxelization(empty.GetImmutableRawData())
        voxelized = voxelizer.VoxelizePointClouds(static, options, wrappers, runtimes.append)
        scenes.check_voxelization(voxelized.GetImmutableRawData())
        assert len(runtimes) == 2 and runtimes[1].RaycastingTime() >= 0.0
        assert messages
        # and exactly the oracle's answer, counts included
        got, counts = voxelizer.VoxelizePointCloudsWithCounts(static, options, wrappers)
        want, want_counts = oracle_voxelize(oracle, scene, scene["clouds"], scene["filter"])
        np.testing.assert_array_equal(counts, want_counts)
        np.testing.assert_array_equal(got.GetImmutableRawData(), want)
    # null logging function must be accepted (:297-311)
    vgt.MakePointCloudVoxelizer({}, None)


def test_random_rays_counts_bit_exact(shared_library, oracle):
    # The 1000 seeded rays of test/voxel_raycasting_test.cpp, each as a one-point cloud whose
    # origin is the cloud pose; plus the at-most-once property on the GPU result.
    g, pairs = scenes.random_ray_pairs()
    dims = tuple(g["voxel_counts"])
    sizes = vgt.VoxelGridSizes.FromVoxelCounts(g["resolution"], dims)
    static = vgt.OccupancyMap(np.eye(4), "world", sizes)
    voxelizer = vgt.B200PointCloudVoxelizer()
    options = vgt.PointCloudVoxelizationFilterOptions()
    batch = 50
    for start in range(0, len(pairs), batch):
        wrappers, expected = [], []
        for origin, point in pairs[start:start + batch]:
            pose = scenes.translation(*origin)
            local = np.asarray(point) - np.asarray(origin)
            wrappers.append(vgt.VectorPointCloudWrapper(local[None, :], pose, g["max_range"]))
            expected.append(oracle.raycast_cloud(local[None, :], pose, g["max_range"], dims,
                                                 g["resolution"]))
        _, counts = voxelizer.VoxelizePointCloudsWithCounts(static, options, wrappers)
        np.testing.assert_array_equal(counts, np.stack(expected))
        assert counts.max() <= 1
        assert not np.any((counts[..., 0] > 0) & (counts[..., 1] > 0))


@pytest.mark.parametrize("filter_options", [(1.0, 1, 1), (0.9, 2, 2)])
def test_camera_scene_counts_and_occupancy(shared_library, oracle, filter_options):
    # BASELINE config 3 at reduced size: clipped rays, NaN pixels, cameras outside the grid.
    scene = synthetic.depth_camera_scene(96, 0.04, 160, 120, max_range=5.0)
    packed = {"static": scene["static_occupancy"], "x_wg": scene["origin_transform"],
              "voxel_size": scene["voxel_size"]}
    static, wrappers = build(packed, scene["clouds"])
    options = vgt.PointCloudVoxelizationFilterOptions(*filter_options)
    got, counts = vgt.B200PointCloudVoxelizer().VoxelizePointCloudsWithCounts(
        static, options, wrappers)
    want, want_counts = oracle_voxelize(oracle, packed, scene["clouds"], filter_options)
    assert want_counts[..., 0].sum() > 100000 and want_counts[..., 1].sum() > 1000
    np.testing.assert_array_equal(counts, want_counts)
    np.testing.assert_array_equal(got.GetImmutableRawData(), want)
    # then the SDF of the voxelized map, end to end
    sdf = got.ExtractSignedDistanceFieldFloat(vgt.SignedDistanceFieldGenerationParameters())
    want_sdf, _ = oracle.sdf(want, scene["voxel_size"])
    np.testing.assert_array_equal(sdf.GetImmutableRawData(), want_sdf)


def test_full_config3_counts(shared_library, oracle):
    # BASELINE config 3 at full size: 4 x 640x480 rays into 256^3.
    scene = synthetic.depth_camera_scene()
    packed = {"static": scene["static_occupancy"], "x_wg": scene["origin_transform"],
              "voxel_size": scene["voxel_size"]}
    static, wrappers = build(packed, scene["clouds"])
    options = vgt.PointCloudVoxelizationFilterOptions(0.9, 2, 2)
    got, counts = vgt.B200PointCloudVoxelizer().VoxelizePointCloudsWithCounts(
        static, options, wrappers)
    want, want_counts = oracle_voxelize(oracle, packed, scene["clouds"], (0.9, 2, 2))
    np.testing.assert_array_equal(counts, want_counts)
    np.testing.assert_array_equal(got.GetImmutableRawData(), want)


def test_argument_errors(shared_library):
    scene = scenes.reference_voxelization_scene()
    static, wrappers = build(scene, scene["clouds"])
    voxelizer = vgt.B200PointCloudVoxelizer()
    options = vgt.PointCloudVoxelizationFilterOptions()
    with pytest.raises(ValueError):       # pcv_if.hpp:281-289
        voxelizer.VoxelizePointClouds(static, options, [wrappers[0], None])
    with pytest.raises(ValueError):       # pcv_if.hpp:30-41
        vgt.PointCloudVoxelizationFilterOptions(0.0, 1, 1)
    with pytest.raises(_capi.BackendUnavailable):   # dev_pcv.hpp:34-46
        vgt.B200PointCloudVoxelizer({"CUDA_DEVICE": 77})


def test_zero_length_rays_with_the_origin_outside_the_grid(shared_library, oracle):
    # A point ON the sensor origin (zero-filled invalid depth return) while the origin is outside
    # the grid: direction 0 / 0. The CPU path marks nothing (cpu_pcv.cpp:229-297: the NaN start
    # index is out of bounds); the device path must not walk from voxel (0, 0, 0).
    dims, voxel = (16, 16, 16), 0.1
    sizes = vgt.VoxelGridSizes.FromVoxelCounts(voxel, dims)
    static = vgt.OccupancyMap(np.eye(4), "world", sizes)
    points = np.zeros((64, 3))
    points[1::2] = [0.3, 0.2, 0.1]          # every other point is a real one
    for origin in ((-1.0, 0.5, 0.5), (0.5, 3.0, 0.5), (-2.0, -2.0, -2.0), (0.5, 0.5, 0.5)):
        pose = scenes.translation(*origin)
        wrapper = vgt.VectorPointCloudWrapper(points, pose, np.inf)
        _, counts = vgt.B200PointCloudVoxelizer().VoxelizePointCloudsWithCounts(
            static, vgt.PointCloudVoxelizationFilterOptions(), [wrapper])
        want = oracle.raycast_cloud(points, pose, np.inf, dims, voxel)
        np.testing.assert_array_equal(counts[0], want)


def test_counts_equal_the_compiled_reference(shared_library):
    # Directly against the REFERENCE'S OWN cpu_pointcloud_voxelization.cpp (oracle/_ref, built in
    # the dev container, travels prebuilt): raw counts and the filtered map, bit for bit.
    from oracle import reference_oracle
    if not reference_oracle.voxelizer_available():
        pytest.skip("oracle/_ref/libvgt_ref.so not built with the voxelizer")
    scene = synthetic.depth_camera_scene(128, 0.04, 320, 240, max_range=4.0)
    packed = {"static": scene["static_occupancy"], "x_wg": scene["origin_transform"],
              "voxel_size": scene["voxel_size"]}
    static, wrappers = build(packed, scene["clouds"])
    x_gw = scenes.inverse_rigid(scene["origin_transform"])
    prepared = [(p, scenes.compose(x_gw, x), r) for p, x, r in scene["clouds"]]
    for filter_options in ((1.0, 1, 1), (0.9, 2, 2)):
        got, counts = vgt.B200PointCloudVoxelizer().VoxelizePointCloudsWithCounts(
            static, vgt.PointCloudVoxelizationFilterOptions(*filter_options), wrappers)
        want, want_counts = reference_oracle.voxelize(scene["static_occupancy"], prepared,
                                                      scene["voxel_size"], *filter_options)
        np.testing.assert_array_equal(counts, want_counts)
        np.testing.assert_array_equal(got.GetImmutableRawData(), want)
    # and the interface call on the posed grid (the reference composes X_GC itself)
    posed = reference_oracle.voxelize_posed(scene["static_occupancy"], scene["origin_transform"],
                                            scene["clouds"], scene["voxel_size"], 0.9, 2, 2)
    np.testing.assert_array_equal(got.GetImmutableRawData(), posed)


def test_float32_clouds_give_the_counts_of_the_same_values_as_doubles(shared_library, oracle):
    # vgt_b200_voxelize_f32: PointCloud2-style float32 points, widened on the device
    scene = synthetic.depth_camera_scene(64, 0.08, 160, 120, max_range=5.0)
    packed = {"static": scene["static_occupancy"], "x_wg": scene["origin_transform"],
              "voxel_size": scene["voxel_size"]}
    rounded = [(p.astype(np.float32), x, r) for p, x, r in scene["clouds"]]
    static, as_double = build(packed, [(p.astype(np.float64), x, r) for p, x, r in rounded])
    as_float = [vgt.Float32PointCloudWrapper(p, x, r) for p, x, r in rounded]
    options = vgt.PointCloudVoxelizationFilterOptions(0.9, 1, 1)
    voxelizer = vgt.B200PointCloudVoxelizer()
    got_f, counts_f = voxelizer.VoxelizePointCloudsWithCounts(static, options, as_float)
    got_d, counts_d = voxelizer.VoxelizePointCloudsWithCounts(static, options, as_double)
    assert counts_f.sum() > 100000
    np.testing.assert_array_equal(counts_f, counts_d)
    np.testing.assert_array_equal(got_f.GetImmutableRawData(), got_d.GetImmutableRawData())
    want, want_counts = oracle_voxelize(
        oracle, packed, [(p.astype(np.float64), x, r) for p, x, r in rounded], (0.9, 1, 1))
    np.testing.assert_array_equal(counts_f, want_counts)
    np.testing.assert_array_equal(got_f.GetImmutableRawData(), want)
