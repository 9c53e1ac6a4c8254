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
