"""CPU: pins the oracle's SDF path against the reference's own known-answer tests
(tests/golden/sdf_generation_test.json <- test/sdf_generation_test.cpp), scipy and brute force."""
import itertools

import numpy as np
import pytest
from scipy import ndimage

from .conftest import occupancy_from_golden_case, random_occupancy


def _case(goldens, name):
    return next(c for c in goldens["cases"] if c["name"] == name)


@pytest.mark.parametrize("threads", [1, 2, 0])
@pytest.mark.parametrize("name", ["FullyFilledTest", "FullyEmptyTest", "CenterObstacleTest",
                                  "CornerObstacleTest", "FaceObstacleTest"])
def test_reference_extrema_and_signs(oracle, sdf_goldens, name, threads):
    # test/sdf_generation_test.cpp:141-259: extrema within kExtremaTolerance, sign matches occupancy.
    case = _case(sdf_goldens, name)
    occupancy, resolution = occupancy_from_golden_case(case)
    assert occupancy.shape == (4, 8, 12)
    tolerance = sdf_goldens["extrema_tolerance"]
    for dtype in (np.float32, np.float64):
        sdf, (lo, hi) = oracle.sdf(occupancy, resolution, threads=threads, dtype=dtype)
        for got, want in zip((lo, hi), case["expected_min_max"]):
            want = float(want)
            assert got == want or abs(got - want) <= tolerance
        assert lo == sdf.min() and hi == sdf.max()
        assert np.all(sdf[occupancy >= 0.5] < 0)
        assert np.all(sdf[occupancy < 0.5] > 0)


@pytest.mark.parametrize("threads", [1, 2])
@pytest.mark.parametrize("name", ["LinearExactTest", "PlanarExactTest", "CubeExactTest"])
def test_reference_exact_cells(oracle, sdf_goldens, name, threads):
    # test/sdf_generation_test.cpp:679-701, 796-902, 997-1055 (EXPECT_FLOAT_EQ = within 4 ulp;
    # we require exact equality, which is stronger).
    case = _case(sdf_goldens, name)
    occupancy, resolution = occupancy_from_golden_case(case)
    sdf, _ = oracle.sdf(occupancy, resolution, threads=threads)
    assert len(case["expected_cells"]) == occupancy.size
    for cell in case["expected_cells"]:
        x, y, z = cell["index"]
        assert sdf[x, y, z] == np.float32(cell["value"]), (name, cell)


def _brute_force_squared(filled: np.ndarray):
    """Independent O(V^2) integer EDT for tiny grids."""
    coords = np.argwhere(np.ones_like(filled, dtype=bool))
    inside = coords[filled.reshape(-1)]
    outside = coords[~filled.reshape(-1)]
    def nearest(targets):
        if len(targets) == 0:
            return np.full(len(coords), np.inf)
        d = ((coords[:, None, :] - targets[None, :, :]) ** 2).sum(axis=2)
        return d.min(axis=1).astype(np.float64)
    return nearest(inside).reshape(filled.shape), nearest(outside).reshape(filled.shape)


@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 1, 7), (3, 1, 9), (5, 7, 9), (9, 10, 11),
                                   (2, 13, 3), (12, 1, 1)])
def test_against_brute_force(oracle, shape):
    rng = np.random.default_rng(hash(shape) % 2 ** 32)
    for fill in (0.0, 0.1, 0.5, 0.9, 1.0):
        occupancy = random_occupancy(rng, shape, fill)
        for unknown_is_filled in (True, False):
            filled = (occupancy > 0.5) | (unknown_is_filled & (occupancy == 0.5))
            to_filled, to_free = oracle.edt_squared(occupancy, unknown_is_filled)
            want_filled, want_free = _brute_force_squared(filled)
            np.testing.assert_array_equal(to_filled, want_filled)
            np.testing.assert_array_equal(to_free, want_free)


def _scipy_sdf(filled: np.ndarray, resolution: float) -> np.ndarray:
    inf = np.float32(np.inf)
    if filled.all():
        return np.full(filled.shape, -inf, dtype=np.float32)
    if not filled.any():
        return np.full(filled.shape, inf, dtype=np.float32)
    outside = ndimage.distance_transform_edt(~filled)
    inside = ndimage.distance_transform_edt(filled)
    return ((outside - inside) * resolution).astype(np.float32)


@pytest.mark.parametrize("shape", [(40, 33, 27), (64, 64, 64), (17, 90, 9)])
def test_against_scipy(oracle, shape):
    # Third opinion. scipy's sqrt happens before the multiply like the reference's, but it rounds
    # through float64 sums differently only in the last ulp, so compare squared distances exactly
    # and the float SDF to 1 ulp.
    rng = np.random.default_rng(5)
    occupancy = random_occupancy(rng, shape, 0.2, blobs=True)
    filled = occupancy >= 0.5
    sdf, _ = oracle.sdf(occupancy, 0.02)
    want = _scipy_sdf(filled, 0.02)
    assert np.all(np.abs(sdf - want) <= np.spacing(np.abs(want)))
    to_filled, to_free = oracle.edt_squared(occupancy)
    np.testing.assert_array_equal(
        to_filled, np.rint(ndimage.distance_transform_edt(~filled) ** 2))
    np.testing.assert_array_equal(
        to_free, np.rint(ndimage.distance_transform_edt(filled) ** 2))


def test_serial_equals_parallel(oracle):
    # test/sdf_generation_test.cpp:1058-1077 runs every case at None() and at N threads.
    rng = np.random.default_rng(11)
    occupancy = random_occupancy(rng, (23, 31, 19), 0.3, blobs=True)
    serial, mm1 = oracle.sdf(occupancy, 0.1, threads=1)
    parallel, mm2 = oracle.sdf(occupancy, 0.1, threads=3)
    np.testing.assert_array_equal(serial, parallel)
    assert mm1 == mm2


def border_formula(occupancy, resolution, unknown_is_filled, sq_filled, sq_free):
    """border mode = min(sq, b^2) in the finalize, b = distance to the
    nearest shell cell over the axes with more than one voxel."""
    shape = occupancy.shape
    border = np.full(shape, np.inf)
    for axis, count in enumerate(shape):
        if count > 1:
            index = np.arange(count, dtype=np.float64)
            along = np.minimum(index + 1, count - index)
            expand = [None, None, None]
            expand[axis] = slice(None)
            border = np.minimum(border, along[tuple(expand)])
    filled = (occupancy > 0.5) | (unknown_is_filled & (occupancy == 0.5))
    squared = np.where(filled, np.minimum(sq_free, border ** 2),
                       np.minimum(sq_filled, border ** 2))
    magnitude = (np.sqrt(squared) * resolution).astype(np.float32)
    return np.where(filled, -magnitude, magnitude)


@pytest.mark.parametrize("shape", [(5, 7, 9), (8, 8, 8), (1, 6, 6), (1, 1, 5), (3, 1, 4),
                                   (1, 1, 1), (12, 9, 10)])
def test_virtual_border_literal_equals_formula(oracle, shape):
    # No reference test covers add_virtual_border; the oracle restates the literal enlarged-grid
    # algorithm (sdfgen.hpp:134-284) and this pins the closed form the CUDA finalize uses.
    rng = np.random.default_rng(3)
    for fill, unknown_is_filled in itertools.product((0.0, 0.1, 0.5, 0.9, 1.0), (True, False)):
        occupancy = random_occupancy(rng, shape, fill)
        literal, _ = oracle.sdf(occupancy, 0.25, unknown_is_filled, add_virtual_border=True)
        sq_filled, sq_free = oracle.edt_squared(occupancy, unknown_is_filled)
        formula = border_formula(occupancy, 0.25, unknown_is_filled, sq_filled, sq_free)
        np.testing.assert_array_equal(literal, formula)


def test_mask_entry_matches_occupancy_entry(oracle):
    rng = np.random.default_rng(9)
    occupancy = random_occupancy(rng, (9, 14, 11), 0.3)
    mask = (occupancy > 0.5) | (occupancy == 0.5)
    a, _ = oracle.sdf(occupancy, 0.5)
    b, _ = oracle.sdf_from_mask(mask, 0.5)
    np.testing.assert_array_equal(a, b)
