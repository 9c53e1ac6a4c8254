"""Scene builders shared by the CPU (oracle) and GPU (C-ABI) tests.

``reference_voxelization_scene`` rebuilds the fixture of the reference's
test/pointcloud_voxelization_test.cpp:160-246 from the literals extracted into
tests/golden/pointcloud_voxelization_test.json."""
import json
from pathlib import Path

import numpy as np

from voxelized_geometry_tools_b200 import synthetic

GOLDEN = Path(__file__).resolve().parent / "golden"


def translation(x, y, z):
    t = np.eye(4)
    t[:3, 3] = (x, y, z)
    return t


def rotation_z(angle):
    c, s = np.cos(angle), np.sin(angle)
    r = np.eye(4)
    r[:2, :2] = [[c, -s], [s, c]]
    return r


def reference_voxelization_scene():
    g = json.loads((GOLDEN / "pointcloud_voxelization_test.json").read_text())
    voxel = g["grid_resolution"]
    shape = tuple(int(np.ceil(s / voxel)) for s in g["grid_size_xyz"])
    x_wg = translation(*g["grid_origin_translation"])
    static = np.zeros(shape, dtype=np.float32)
    static[:, :, 0] = 1.0  # :183-188 bottom cells filled
    x_co = synthetic.optical_from_physical()  # :192-194
    lo, hi, step = g["lattice_min_max_step"]
    lattice = np.arange(lo, hi + 0.5 * step, step)  # 129 exact binary fractions
    xs, ys = np.meshgrid(lattice, lattice, indexing="ij")
    xs, ys = xs.reshape(-1), ys.reshape(-1)
    near, far = g["near_depth"], g["far_depth"]
    cloud1 = np.stack([xs, ys, np.where(xs <= 0.0, near, far)], axis=1)  # :201-209
    cloud2 = np.stack([xs, ys, np.where(xs >= 0.0, near, far)], axis=1)  # :219-227
    x_wc1 = translation(*g["camera1_translation"]) @ x_co
    x_wc2 = translation(*g["camera2_translation"]) @ rotation_z(np.pi / 2) @ x_co  # :212-215
    x_wc3 = x_co.copy()  # :230-235, empty cloud
    clouds = [(cloud1, x_wc1, np.inf), (cloud2, x_wc2, np.inf),
              (np.zeros((0, 3)), x_wc3, np.inf)]
    filter_options = (g["percent_seen_free"], g["outlier_points_threshold"],
                      g["num_cameras_seen_free"])
    return {"static": static, "x_wg": x_wg, "voxel_size": voxel, "clouds": clouds,
            "filter": filter_options}


def inverse_rigid(transform):
    """The front end's fixed-order inverse (grids.inverse_rigid)."""
    from voxelized_geometry_tools_b200.grids import inverse_rigid as fixed_order_inverse
    return fixed_order_inverse(transform)


def compose(a, b):
    """a * b in the fixed operation order every front end uses for X_GC (grids.compose_rigid)."""
    from voxelized_geometry_tools_b200.grids import compose_rigid
    return compose_rigid(a, b)


def check_empty_voxelization(occupancy):
    """test/pointcloud_voxelization_test.cpp:84-111."""
    assert np.all(occupancy[:, :, 0] == 1.0)
    assert np.all(occupancy[:, :, 1:] == 0.5)


def check_voxelization(occupancy):
    """test/pointcloud_voxelization_test.cpp:113-158."""
    assert np.all(occupancy[:, :, 0] == 1.0)
    assert np.all(occupancy[3, 3:, 1:] == 0.0)   # seen empty
    assert np.all(occupancy[3:, 3, 1:] == 0.0)
    assert np.all(occupancy[4, 4:, 1:] == 1.0)   # seen filled
    assert np.all(occupancy[4:, 4, 1:] == 1.0)
    assert np.all(occupancy[5:, 5:, 1:] == 0.5)  # shadowed


def random_ray_pairs(count=1000):
    """test/voxel_raycasting_test.cpp:36-100: mt19937_64(42), generate_canonical, [-2, 7]^3."""
    g = json.loads((GOLDEN / "voxel_raycasting_test.json").read_text())
    rng = synthetic.Mt19937_64(g["seed"])
    lo, hi = g["min_axis_value"], g["max_axis_value"]

    def sample():
        # common_robotics_utilities::math::Interpolate(first, second, ratio)
        return [(hi * r) + (lo * (1.0 - r)) for r in (rng.canonical() for _ in range(3))]

    pairs = [(sample(), sample()) for _ in range(count)]
    return g, pairs
