"""GPU: the other map types' SDF entry points through the C-ABI
(vgt_b200_sdf_from_cells_*, vgt_b200_sdf_per_object_*, vgt_b200_sdf_free_and_named_*) against the
oracle. Bar: float / double SDF and min/max bit-exact."""
import numpy as np
import pytest

import voxelized_geometry_tools_b200 as vgt
from voxelized_geometry_tools_b200 import _capi, grids

from .conftest import occupancy_from_golden_case, random_occupancy

pytestmark = pytest.mark.gpu

IDENTITY = np.eye(4)


def params(unknown_is_filled=True, add_virtual_border=False):
    return vgt.SignedDistanceFieldGenerationParameters(
        float("inf"), None, unknown_is_filled, add_virtual_border)


def sizes_of(shape, resolution):
    return vgt.VoxelGridSizes.FromVoxelCounts(resolution, shape)


def random_cells(rng, shape, cell_dtype, num_objects=6):
    occupancy = random_occupancy(rng, shape, 0.15, blobs=True)
    cells = np.zeros(shape, dtype=cell_dtype)
    cells["occupancy"] = occupancy
    # objects as coarse blocks so that every id forms real shapes; id 0 = unnamed
    block = np.add.outer(np.add.outer(np.arange(shape[0]) // 5, np.arange(shape[1]) // 7),
                         np.arange(shape[2]) // 6)
    if "object_id" in cell_dtype.names:
        cells["object_id"] = (block * 2654435761 % num_objects).astype(np.uint32)
    if "component" in cell_dtype.names:
        cells["component"] = rng.integers(0, 1000, size=shape, dtype=np.uint32)
    return cells


def assert_same(sdf, want, extrema):
    np.testing.assert_array_equal(sdf.GetImmutableRawData(), want)
    assert sdf.IsLocked()
    assert sdf.GetMinimumMaximum() == extrema


def test_four_map_types_agree_on_the_reference_cases(shared_library, sdf_goldens, oracle):
    # test/sdf_generation_test.cpp:152-189: every map type yields the OccupancyMap's SDF.
    for case in sdf_goldens["cases"]:
        occupancy, resolution = occupancy_from_golden_case(case)
        sizes = sizes_of(occupancy.shape, resolution)
        want, extrema = oracle.sdf(occupancy, resolution)
        reference = vgt.OccupancyMap(IDENTITY, "f", sizes, data=occupancy) \
            .ExtractSignedDistanceFieldFloat(params())
        assert_same(reference, want, extrema)
        component = np.zeros(occupancy.shape, dtype=grids.OCCUPANCY_COMPONENT_CELL)
        component["occupancy"] = occupancy
        assert_same(vgt.OccupancyComponentMap(IDENTITY, "f", sizes, component)
                    .ExtractSignedDistanceFieldFloat(params()), want, extrema)
        for map_type in (vgt.TaggedObjectOccupancyMap, vgt.TaggedObjectOccupancyComponentMap):
            cells = np.zeros(occupancy.shape, dtype=map_type.CELL)
            cells["occupancy"] = occupancy
            cells["object_id"] = 1
            assert_same(map_type(IDENTITY, "f", sizes, cells)
                        .ExtractSignedDistanceFieldFloat([], params()), want, extrema)


@pytest.mark.parametrize("map_type", [vgt.TaggedObjectOccupancyMap,
                                      vgt.TaggedObjectOccupancyComponentMap])
def test_objects_to_use(shared_library, oracle, map_type):
    rng = np.random.default_rng(51)
    shape = (23, 31, 40)
    cells = random_cells(rng, shape, map_type.CELL)
    tagged = map_type(IDENTITY, "f", sizes_of(shape, 0.04), cells)
    for objects, unknown, border, dtype in (
            ([], True, False, np.float32), ([3], True, False, np.float32),
            ([1, 4, 5, 4], False, True, np.float32), ([77], True, False, np.float32),
            ([0, 2], True, True, np.float64)):
        want, extrema = oracle.sdf_from_cells(cells, 0.04, unknown, border, objects, dtype=dtype)
        got = tagged.ExtractSignedDistanceField(objects, params(unknown, border), dtype)
        assert got.GetImmutableRawData().dtype == np.dtype(dtype)
        assert_same(got, want, extrema)


def test_per_object_batches(shared_library, oracle):
    rng = np.random.default_rng(52)
    shape = (20, 18, 36)
    cells = random_cells(rng, shape, grids.TAGGED_OBJECT_OCCUPANCY_CELL, num_objects=5)
    tagged = vgt.TaggedObjectOccupancyMap(IDENTITY, "f", sizes_of(shape, 0.1), cells)
    separate = tagged.MakeSeparateObjectSDFs([4, 1, 9], params())
    assert sorted(separate) == [1, 4, 9]
    for object_id, sdf in separate.items():
        want, extrema = oracle.sdf_from_cells(cells, 0.1, objects_to_use=[object_id])
        assert_same(sdf, want, extrema)
    everything = tagged.MakeAllObjectSDFs(params(), np.float64)
    present = sorted(int(i) for i in np.unique(cells["object_id"]) if i > 0)
    assert sorted(everything) == present
    for object_id, sdf in everything.items():
        want, extrema = oracle.sdf_from_cells(cells, 0.1, objects_to_use=[object_id],
                                              dtype=np.float64)
        assert_same(sdf, want, extrema)
    assert tagged.MakeSeparateObjectSDFs([], params()) == {}


@pytest.mark.parametrize("map_type", [vgt.TaggedObjectOccupancyMap,
                                      vgt.TaggedObjectOccupancyComponentMap])
def test_free_and_named_objects(shared_library, oracle, map_type):
    rng = np.random.default_rng(53)
    shape = (26, 22, 34)
    cells = random_cells(rng, shape, map_type.CELL, num_objects=3)
    tagged = map_type(IDENTITY, "f", sizes_of(shape, 0.05), cells)
    for unknown, border, dtype in ((True, False, np.float32), (False, True, np.float64)):
        want, extrema = oracle.sdf_free_and_named(cells, 0.05, unknown, border, dtype)
        got = tagged.ExtractFreeAndNamedObjectsSignedDistanceField(params(unknown, border), dtype)
        assert_same(got, want, extrema)


def test_tagged_map_equals_the_references_own_members(shared_library):
    # the device entries against the reference's own TaggedObjectOccupancyMap members
    # (tagged_object_occupancy_map.hpp / .cpp compiled unmodified, oracle/_ref/libvgt_ref_maps.so)
    from oracle import reference_oracle
    if not reference_oracle.maps_available():
        pytest.skip("oracle/_ref/libvgt_ref_maps.so not built")
    rng = np.random.default_rng(54)
    shape = (21, 26, 38)
    cells = random_cells(rng, shape, grids.TAGGED_OBJECT_OCCUPANCY_CELL, num_objects=5)
    tagged = vgt.TaggedObjectOccupancyMap(IDENTITY, "f", sizes_of(shape, 0.04), cells)
    for objects, unknown, border, dtype in (
            ([], True, False, np.float32), ([2, 4], False, True, np.float32),
            ([3], True, True, np.float64)):
        want, extrema = reference_oracle.tagged_map_sdf(cells, 0.04, objects, unknown, border,
                                                        dtype=dtype)
        assert_same(tagged.ExtractSignedDistanceField(objects, params(unknown, border), dtype),
                    want, extrema)
    for unknown, border, dtype in ((True, False, np.float32), (False, True, np.float64)):
        want, extrema = reference_oracle.tagged_map_sdf(cells, 0.04, (), unknown, border,
                                                        free_and_named=True, dtype=dtype)
        assert_same(tagged.ExtractFreeAndNamedObjectsSignedDistanceField(params(unknown, border),
                                                                         dtype), want, extrema)


def test_cell_entry_rejects_bad_arguments(shared_library):
    lib = _capi.library()
    cells = np.zeros((2, 2, 2, 3), dtype=np.uint32)
    out = np.zeros((2, 2, 2), dtype=np.float32)
    assert lib.vgt_b200_sdf_from_cells_f32(cells.ctypes.data, 12, 2, 2, 2, 1.0, 1, 0, None, 0, 0,
                                           out.ctypes.data, None, None) \
        == _capi.ERR_INVALID_ARGUMENT
    assert "cell size" in _capi.last_error()
    assert lib.vgt_b200_sdf_from_cells_f32(cells.ctypes.data, 8, 2, 2, 2, 1.0, 1, 0, None, 3, 0,
                                           out.ctypes.data, None, None) \
        == _capi.ERR_INVALID_ARGUMENT
