"""CPU: the oracle's restatement of the other map types' SDF entry points (SURVEY.md section 8f,
rank 1) against the reference's own assertions: the four map types produce the same SDF for the
same occupancy (test/sdf_generation_test.cpp:152-189), and hand-checked predicate / merge cases."""
import numpy as np
import pytest

from oracle import reference_oracle
from voxelized_geometry_tools_b200 import grids

from .conftest import occupancy_from_golden_case, random_occupancy


def tagged_cells(occupancy, object_ids, dtype=grids.TAGGED_OBJECT_OCCUPANCY_CELL):
    cells = np.zeros(occupancy.shape, dtype=dtype)
    cells["occupancy"] = occupancy
    cells["object_id"] = object_ids
    return cells


def test_four_map_types_agree_on_the_reference_cases(oracle, sdf_goldens):
    for case in sdf_goldens["cases"]:
        occupancy, resolution = occupancy_from_golden_case(case)
        want, want_extrema = oracle.sdf(occupancy, resolution)
        component = np.zeros(occupancy.shape, dtype=grids.OCCUPANCY_COMPONENT_CELL)
        component["occupancy"] = occupancy
        component["component"] = 7
        ids = np.arange(occupancy.size, dtype=np.uint32).reshape(occupancy.shape) % 5
        for cells in (component, tagged_cells(occupancy, ids),
                      tagged_cells(occupancy, ids, grids.TAGGED_OBJECT_OCCUPANCY_COMPONENT_CELL)):
            objects = () if "object_id" in (cells.dtype.names or ()) else ()
            got, extrema = oracle.sdf_from_cells(cells, resolution, objects_to_use=objects)
            np.testing.assert_array_equal(got, want)
            assert extrema == want_extrema


def test_object_predicates(oracle):
    occupancy = np.array([0.0, 1.0, 0.5, 1.0, 0.7, 0.2], dtype=np.float32).reshape(1, 1, 6)
    ids = np.array([0, 3, 3, 0, 9, 9], dtype=np.uint32).reshape(1, 1, 6)
    cells = tagged_cells(occupancy, ids)
    mask = oracle.cells_filled_mask
    assert mask(cells).ravel().tolist() == [False, True, True, True, True, False]
    assert mask(cells, unknown_is_filled=False).ravel().tolist() == [False, True, False, True, True, False]
    assert mask(cells, objects_to_use=[3]).ravel().tolist() == [False, True, True, False, False, False]
    assert mask(cells, objects_to_use=[9, 3, 3]).ravel().tolist() == [False, True, True, False, True, False]
    assert mask(cells, objects_to_use=[5]).ravel().tolist() == [False] * 6
    assert mask(cells, named_only=True).ravel().tolist() == [False, True, True, False, True, False]


def test_free_and_named_merge(oracle):
    # x line: free | unnamed obstacle (id 0) | free | named obstacle (id 2)
    occupancy = np.array([0, 0, 1, 1, 0, 0, 1], dtype=np.float32).reshape(7, 1, 1)
    ids = np.array([0, 0, 0, 0, 0, 0, 2], dtype=np.uint32).reshape(7, 1, 1)
    got, (lo, hi) = oracle.sdf_free_and_named(tagged_cells(occupancy, ids), 1.0)
    # free cells keep the distance to the nearest obstacle of any kind; the unnamed obstacle is
    # not inside a named object, so it reads 0; the named cell reads its depth inside named space
    assert got.ravel().tolist() == [2.0, 1.0, 0.0, 0.0, 1.0, 1.0, -1.0]
    assert (lo, hi) == (-1.0, 2.0)


@pytest.mark.skipif(not reference_oracle.maps_available(),
                    reason="oracle/_ref/libvgt_ref_maps.so not built")
def test_restatement_equals_the_references_own_tagged_map_members(oracle):
    # TaggedObjectOccupancyMap::ExtractSignedDistanceField<T>(objects_to_use, ...) and
    # ::ExtractFreeAndNamedObjectsSignedDistanceField<T> on the reference's own class (its
    # tagged_object_occupancy_map.hpp / .cpp compiled unmodified): predicate with and without
    # ids, ids that do not occur, repeated ids, both unknown rules, border, both scalar types.
    rng = np.random.default_rng(61)
    for shape in ((9, 7, 11), (14, 20, 6)):
        occupancy = random_occupancy(rng, shape, 0.25, unknown=0.15)
        ids = rng.integers(0, 5, size=shape).astype(np.uint32)
        cells = tagged_cells(occupancy, ids)
        for objects, unknown, border, dtype in (
                ((), True, False, np.float32), ((3,), True, False, np.float32),
                ((1, 4, 4), False, True, np.float32), ((77,), True, False, np.float64),
                ((0, 2), False, True, np.float64)):
            mine, mine_extrema = oracle.sdf_from_cells(cells, 0.2, unknown, border, objects,
                                                       dtype=dtype)
            theirs, their_extrema = reference_oracle.tagged_map_sdf(
                cells, 0.2, objects, unknown, border, dtype=dtype)
            np.testing.assert_array_equal(mine, theirs)
            assert mine_extrema == their_extrema
        for unknown, border, dtype in ((True, False, np.float32), (False, True, np.float64)):
            mine, mine_extrema = oracle.sdf_free_and_named(cells, 0.2, unknown, border, dtype)
            theirs, their_extrema = reference_oracle.tagged_map_sdf(
                cells, 0.2, (), unknown, border, free_and_named=True, dtype=dtype)
            np.testing.assert_array_equal(mine, theirs)
            assert mine_extrema == their_extrema


@pytest.mark.skipif(not reference_oracle.maps_available(),
                    reason="oracle/_ref/libvgt_ref_maps.so not built")
def test_restatement_equals_the_references_own_component_map_members(oracle):
    # OccupancyComponentMap and TaggedObjectOccupancyComponentMap (headers and sources
    # unmodified): their SDF members against the restatement, component / segment words filled
    # with noise (they must not matter).
    rng = np.random.default_rng(62)
    shape = (10, 9, 13)
    occupancy = random_occupancy(rng, shape, 0.3, unknown=0.15)
    component = np.zeros(shape, dtype=grids.OCCUPANCY_COMPONENT_CELL)
    component["occupancy"] = occupancy
    component["component"] = rng.integers(0, 2 ** 32, size=shape, dtype=np.uint64).astype(np.uint32)
    for unknown, border, dtype in ((True, False, np.float32), (False, True, np.float64)):
        mine, mine_extrema = oracle.sdf_from_cells(component, 0.1, unknown, border, dtype=dtype)
        theirs, their_extrema = reference_oracle.component_map_sdf(
            component, 0.1, (), unknown, border, dtype=dtype)
        np.testing.assert_array_equal(mine, theirs)
        assert mine_extrema == their_extrema
    tagged = tagged_cells(occupancy, rng.integers(0, 4, size=shape).astype(np.uint32),
                          grids.TAGGED_OBJECT_OCCUPANCY_COMPONENT_CELL)
    tagged["component"] = rng.integers(0, 1000, size=shape).astype(np.uint32)
    tagged["spatial_segment"] = rng.integers(0, 1000, size=shape).astype(np.uint32)
    for objects, unknown, border, dtype in (((), True, False, np.float32),
                                            ((2, 3), False, True, np.float32),
                                            ((1,), True, True, np.float64)):
        mine, mine_extrema = oracle.sdf_from_cells(tagged, 0.1, unknown, border, objects,
                                                   dtype=dtype)
        theirs, their_extrema = reference_oracle.component_map_sdf(
            tagged, 0.1, objects, unknown, border, dtype=dtype)
        np.testing.assert_array_equal(mine, theirs)
        assert mine_extrema == their_extrema
    for unknown, border, dtype in ((True, False, np.float32), (False, True, np.float64)):
        mine, mine_extrema = oracle.sdf_free_and_named(tagged, 0.1, unknown, border, dtype)
        theirs, their_extrema = reference_oracle.component_map_sdf(
            tagged, 0.1, (), unknown, border, free_and_named=True, dtype=dtype)
        np.testing.assert_array_equal(mine, theirs)
        assert mine_extrema == their_extrema
