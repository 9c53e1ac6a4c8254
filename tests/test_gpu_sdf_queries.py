"""GPU: the batched SignedDistanceField queries (csrc/sdf_queries.cu through the C-ABI) against
the statement-by-statement restatement in oracle/sdf_queries_oracle.py: bit-exact, statuses
included. (Against the reference itself the trilinear blend is parity-unpinned - it lives in an
unvendored dependency - with 1e-12 relative as the stated tolerance; see the oracle's header.)"""
import numpy as np
import pytest

from voxelized_geometry_tools_b200 import synthetic

pytestmark = pytest.mark.gpu


def _posed(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    m = np.eye(4)
    m[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                 [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                 [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    m[:3, 3] = rng.uniform(-2, 2, 3)
    return m


def _world_points(rng, pose, dims, res, count):
    """Points in and slightly around the grid, some exactly on cell centres and faces."""
    extent = np.array(dims) * res
    grid = rng.uniform(-0.15, 1.15, (count, 3)) * extent
    grid[: count // 8] = (rng.integers(0, 6, (count // 8, 3)) + 0.5) * res     # cell centres
    grid[count // 8: count // 4] = rng.integers(0, 6, (count // 8, 3)) * res   # cell corners
    return grid @ pose[:3, :3].T + pose[:3, 3]


@pytest.fixture(scope="module")
def scene(shared_library):
    import torch
    from voxelized_geometry_tools_b200 import device as vdev
    from oracle import sdf_queries_oracle
    rng = np.random.default_rng(11)
    dims, res = (28, 22, 25), 0.05
    occupancy = synthetic.clustered_spheres_occupancy(dims)
    dev = torch.device("cuda", 0)
    sdf, _ = vdev.signed_distance_field(torch.from_numpy(occupancy).to(dev), res)
    pose = _posed(rng)
    device_sdf = vdev.DeviceSignedDistanceField(sdf, res, pose)
    checker = sdf_queries_oracle.SdfOracle(sdf.cpu().numpy(), res, pose)
    points = _world_points(rng, pose, dims, res, 3000)
    return {"torch": torch, "dev": dev, "gpu": device_sdf, "oracle": checker, "points": points,
            "res": res, "dims": dims, "pose": pose}


def _compare(got_values, got_valid, want):
    want_valid = np.array([w[0] for w in want], dtype=np.uint8)
    np.testing.assert_array_equal(got_valid, want_valid)
    want_values = np.array([w[1] for w in want], dtype=np.float64)
    np.testing.assert_array_equal(got_values.reshape(want_values.shape), want_values)


def test_estimate_location_distance(scene):
    points = scene["torch"].from_numpy(scene["points"]).to(scene["dev"])
    values, valid = scene["gpu"].EstimateLocationDistance(points)
    want = [scene["oracle"].estimate_distance(p) for p in scene["points"]]
    _compare(values.cpu().numpy(), valid.cpu().numpy(), want)
    assert 0.5 < np.mean([w[0] for w in want]) < 1.0    # inside and outside points both occur


@pytest.mark.parametrize("edge", [False, True])
def test_coarse_gradient(scene, edge):
    points = scene["torch"].from_numpy(scene["points"]).to(scene["dev"])
    values, valid = scene["gpu"].GetLocationCoarseGradient(points, edge)
    want = [scene["oracle"].coarse_gradient(p, edge) for p in scene["points"]]
    _compare(values.cpu().numpy(), valid.cpu().numpy(), want)


@pytest.mark.parametrize("window", [0.05, 0.013, 0.4])
def test_fine_gradient(scene, window):
    points = scene["torch"].from_numpy(scene["points"][:1500]).to(scene["dev"])
    values, valid = scene["gpu"].GetLocationFineGradient(points, window)
    want = [scene["oracle"].fine_gradient(p, window) for p in scene["points"][:1500]]
    _compare(values.cpu().numpy(), valid.cpu().numpy(), want)


@pytest.mark.parametrize("minimum_distance,multiplier", [(0.0, 0.1), (0.04, 0.25)])
def test_project_out_of_collision(scene, minimum_distance, multiplier):
    subset = scene["points"][:600]
    points = scene["torch"].from_numpy(subset).to(scene["dev"])
    values, valid = scene["gpu"].ProjectLocationOutOfCollisionToMinimumDistance(
        points, minimum_distance, multiplier, max_steps=20000)
    want = [scene["oracle"].project_out_of_collision(p, minimum_distance, multiplier, 20000)
            for p in subset]
    _compare(values.cpu().numpy(), valid.cpu().numpy(), want)
    moved = np.any(values.cpu().numpy() != subset, axis=1) & (valid.cpu().numpy() == 1)
    assert moved.sum() > 20          # points inside obstacles were pushed out
    # and a projected point is out of collision
    again, ok = scene["gpu"].EstimateLocationDistance(values[valid == 1])
    assert bool((again[ok == 1] > minimum_distance).all())


def test_query_properties_on_a_flat_wall(shared_library):
    # x < 8 filled: away from the wall the estimate is x-distance to the wall surface, the
    # gradient +x (rotated by the pose), and projection moves points along +x only.
    import torch
    from voxelized_geometry_tools_b200 import device as vdev
    dev = torch.device("cuda", 0)
    occupancy = np.zeros((32, 16, 16), dtype=np.float32)
    occupancy[:8] = 1.0
    res = 0.1
    sdf, _ = vdev.signed_distance_field(torch.from_numpy(occupancy).to(dev), res)
    field = vdev.DeviceSignedDistanceField(sdf, res)
    points = torch.tensor([[1.55, 0.8, 0.8], [2.0, 0.5, 1.0], [0.35, 0.8, 0.8]],
                          dtype=torch.float64, device=dev)
    distance, valid = field.EstimateLocationDistance(points)
    assert valid.tolist() == [1, 1, 1]
    np.testing.assert_allclose(distance.cpu().numpy(), [0.75, 1.2, -0.45], rtol=0, atol=1e-6)
    gradient, valid = field.GetLocationCoarseGradient(points)
    np.testing.assert_allclose(gradient.cpu().numpy(), [[1, 0, 0]] * 3, rtol=0, atol=1e-6)
    projected, valid = field.ProjectLocationOutOfCollision(points)
    assert valid.tolist() == [1, 1, 1]
    out = projected.cpu().numpy()
    np.testing.assert_array_equal(out[:2], points.cpu().numpy()[:2])     # already free
    assert out[2, 0] > 0.8 and out[2, 1] == 0.8 and out[2, 2] == 0.8
    with pytest.raises(ValueError):
        field.EstimateLocationDistance(torch.zeros((4, 2), dtype=torch.float64, device=dev))


def test_local_extrema_map_of_an_sdf(scene):
    got = scene["gpu"].ComputeLocalExtremaMap().cpu().numpy()
    want = scene["oracle"].local_extrema_map()
    np.testing.assert_array_equal(got, want)
    assert np.isinf(got).any() and np.isfinite(got).any()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_local_extrema_map_with_loops(shared_library, seed):
    """A rough random field: plenty of walks that close loops (two cells pointing at each other
    and longer ones), where the result depends on the order the reference starts its walks in."""
    import torch
    from voxelized_geometry_tools_b200 import device as vdev
    from oracle import sdf_queries_oracle
    rng = np.random.default_rng(seed)
    dims, res = (14, 11, 13), 0.1
    field = rng.normal(size=dims).astype(np.float32) * 0.3
    field[rng.random(dims) < 0.1] = 0.0
    if seed == 2:
        field[2:5, 2:5, 2:5] = np.inf      # inf - inf = NaN gradients: cells that stay put
    pose = _posed(rng) if seed else np.eye(4)
    checker = sdf_queries_oracle.SdfOracle(field, res, pose)
    # the field has loops (otherwise this test checks nothing new)
    loops = 0
    for start in np.ndindex(*dims):
        seen, current = {start}, start
        while True:
            g = checker.coarse_gradient_at_index(*current, True)[1]
            if checker._effectively_flat(g):
                break
            current = checker._next_from_gradient(current, g)
            if not checker._in_bounds(current):
                break
            if current in seen:
                loops += 1
                break
            seen.add(current)
    assert loops > 20
    dev = torch.device("cuda", 0)
    device_sdf = vdev.DeviceSignedDistanceField(torch.from_numpy(field).to(dev), res, pose)
    got = device_sdf.ComputeLocalExtremaMap().cpu().numpy()
    np.testing.assert_array_equal(got, checker.local_extrema_map())
