"""CPU: oracle/sdf_queries_oracle.py (the restatement the device kernels are checked against)
against the reference's OWN SignedDistanceField class, compiled unmodified into oracle/_ref
(ref_shim/ref_sdf_queries_entry.cpp): bit for bit, statuses and exceptions included. The
trilinear blend inside EstimateLocationDistance is the one piece that comes from an unvendored
dependency (restated in ref_shim/common_robotics_utilities/math.hpp); everything around it is the
reference's code running as written."""
import numpy as np
import pytest

from oracle import oracle as restated, reference_oracle, sdf_queries_oracle
from voxelized_geometry_tools_b200 import synthetic

pytestmark = pytest.mark.skipif(
    not reference_oracle.available(),
    reason="oracle/_ref was not built (needs /root/reference at build time)")


def _pose(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    m = np.eye(4)
    m[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                 [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                 [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    m[:3, 3] = rng.uniform(-2, 2, 3)
    return m


def _points(rng, pose, dims, res, count):
    extent = np.array(dims) * res
    grid = rng.uniform(-0.15, 1.15, (count, 3)) * extent
    grid[: count // 8] = (rng.integers(0, 6, (count // 8, 3)) + 0.5) * res     # cell centres
    grid[count // 8: count // 4] = rng.integers(0, 6, (count // 8, 3)) * res   # cell corners
    return grid @ pose[:3, :3].T + pose[:3, 3]


@pytest.fixture(scope="module", params=[0, 1])
def case(request):
    rng = np.random.default_rng(100 + request.param)
    dims, res = (20, 17, 19), 0.05
    field, _ = restated.sdf(synthetic.clustered_spheres_occupancy(dims), res)
    pose = _pose(rng) if request.param else np.eye(4)
    return {"field": field, "res": res, "pose": pose, "dims": dims,
            "checker": sdf_queries_oracle.SdfOracle(field, res, pose),
            "points": _points(rng, pose, dims, res, 1200)}


def _assert_equal(case, kind, a, b, restated_call):
    values, status = reference_oracle.sdf_query(case["field"], case["res"], case["pose"], kind,
                                                case["points"], a, b)
    want = [restated_call(p) for p in case["points"]]
    np.testing.assert_array_equal(status, np.array([w[0] for w in want], dtype=np.uint8))
    np.testing.assert_array_equal(
        values, np.array([np.atleast_1d(w[1]) for w in want], dtype=np.float64))
    return status


def test_estimate_distance(case):
    status = _assert_equal(case, reference_oracle.QUERY_ESTIMATE_DISTANCE, 0.0, 0.0,
                           case["checker"].estimate_distance)
    assert (status == 1).any() and (status == 0).any()


@pytest.mark.parametrize("edge", [False, True])
def test_coarse_gradient(case, edge):
    _assert_equal(case, reference_oracle.QUERY_COARSE_GRADIENT, float(edge), 0.0,
                  lambda p: case["checker"].coarse_gradient(p, edge))


@pytest.mark.parametrize("window", [0.05, 0.13, 5.0])
def test_fine_gradient(case, window):
    status = _assert_equal(case, reference_oracle.QUERY_FINE_GRADIENT, window, 0.0,
                           lambda p: case["checker"].fine_gradient(p, window))
    if window == 5.0:
        assert (status == 2).any()      # the reference throws when the window leaves the grid


@pytest.mark.parametrize("minimum_distance,multiplier", [(0.0, 0.1), (0.04, 0.25)])
def test_project_out_of_collision(case, minimum_distance, multiplier):
    # The reference's loop has no step limit (a point that oscillates between two cells never
    # returns), so only the points the restatement finishes within 5000 steps go to it.
    subset = case["points"][:500]
    want = [case["checker"].project_out_of_collision(p, minimum_distance, multiplier, 5000)
            for p in subset]
    finishes = np.array([w[0] != sdf_queries_oracle.THROWS for w in want])
    assert finishes.sum() > 400
    values, status = reference_oracle.sdf_query(
        case["field"], case["res"], case["pose"], reference_oracle.QUERY_PROJECT,
        subset[finishes], minimum_distance, multiplier)
    kept = [w for w, ok in zip(want, finishes) if ok]
    np.testing.assert_array_equal(status, np.array([w[0] for w in kept], dtype=np.uint8))
    np.testing.assert_array_equal(values, np.array([w[1] for w in kept], dtype=np.float64))
    assert np.any(values != subset[finishes])       # points inside obstacles were moved


def test_local_extrema_map_of_an_sdf(case):
    got = case["checker"].local_extrema_map()
    want = reference_oracle.sdf_local_extrema_map(case["field"], case["res"], case["pose"])
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_local_extrema_map_of_a_rough_field_with_loops(seed):
    rng = np.random.default_rng(seed)
    dims = (14, 11, 13)
    field = rng.normal(size=dims).astype(np.float32) * 0.3
    field[rng.random(dims) < 0.1] = 0.0
    if seed == 2:
        field[2:5, 2:5, 2:5] = np.inf
    pose = _pose(rng) if seed else np.eye(4)
    got = sdf_queries_oracle.SdfOracle(field, 0.1, pose).local_extrema_map()
    want = reference_oracle.sdf_local_extrema_map(field, 0.1, pose)
    np.testing.assert_array_equal(got, want)
