"""The reference's grid files (SDFZ / SDFR, CMGZ / CMGR): product (C-ABI, host functions: no
device needed) against the restated oracle and against the reference's own SaveToFile /
LoadFromFile members compiled over oracle/ref_shim. See oracle/grid_files_oracle.py for what is
pinned (everything the reference tree owns) and what is not (the third-party grid layout)."""
import ctypes
import zlib

import numpy as np
import pytest

from oracle import grid_files_oracle as gfo
from oracle import reference_oracle
from voxelized_geometry_tools_b200 import grid_files, grids


@pytest.fixture(scope="module")
def reference_library():
    """oracle/_ref/libvgt_ref.so with the file members (built where /root/reference exists)."""
    if not reference_oracle.available():
        pytest.skip("oracle/_ref is not built (needs /root/reference)")
    handle = ctypes.CDLL(str(reference_oracle._PATH))
    if not hasattr(handle, "vgt_ref_sdf_save_to_file"):
        pytest.skip("oracle/_ref predates the file members: make -C oracle ref")
    return handle


def origin_transform(rng):
    # a proper rotation (QR of a random matrix) and a translation, row-major 4x4
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    transform = np.eye(4)
    transform[:3, :3] = q
    transform[:3, 3] = rng.uniform(-3, 3, size=3)
    return transform


def make_sdf(rng, dtype, shape=(5, 4, 7), locked=True, oob=np.inf):
    data = rng.normal(size=shape).astype(dtype)
    data[0, 0, 0] = -np.inf if dtype == np.float32 else 1e300
    sizes = grids.VoxelGridSizes.FromVoxelCounts(0.25, shape)
    sdf = grids.SignedDistanceField(origin_transform(rng), "some_frame", sizes, data, oob)
    if locked:
        sdf.Lock()
    return sdf


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("compress", [False, True])
@pytest.mark.parametrize("locked", [False, True])
def test_sdf_file_equals_the_oracle_and_round_trips(tmp_path, dtype, compress, locked):
    rng = np.random.default_rng(7)
    sdf = make_sdf(rng, dtype, locked=locked)
    path = tmp_path / "field.sdf"
    grid_files.SaveSignedDistanceFieldToFile(sdf, path, compress)
    content = path.read_bytes()
    assert content[:4] == (b"SDFZ" if compress else b"SDFR")
    payload = zlib.decompress(content[4:]) if compress else content[4:]
    want = gfo.serialize_grid(sdf.GetImmutableRawData(), 0.25,
                              sdf.OriginTransform().T.reshape(-1), np.inf, np.inf, "some_frame",
                              locked=locked)
    assert payload == want
    # the oracle writes the same file (same zlib, same level), and reads ours
    other = tmp_path / "oracle.sdf"
    gfo.save_to_file(other, gfo.SDF_MAGIC, want, compress)
    assert other.read_bytes() == content
    parsed = gfo.deserialize_grid(gfo.load_from_file(path, gfo.SDF_MAGIC), dtype, True)
    assert parsed["locked"] == locked and parsed["frame"] == "some_frame"
    assert np.array_equal(parsed["cells"], sdf.GetImmutableRawData())
    assert parsed["bytes_read"] == len(want)
    # and the product reads the oracle's
    loaded = grid_files.LoadSignedDistanceFieldFromFile(other, dtype)
    assert loaded.IsLocked() == locked and loaded.Frame() == "some_frame"
    assert loaded.GetImmutableRawData().dtype == dtype
    assert np.array_equal(loaded.GetImmutableRawData(), sdf.GetImmutableRawData())
    assert np.array_equal(loaded.OriginTransform(), sdf.OriginTransform())
    assert loaded.Resolution() == 0.25 and loaded._oob_value == np.inf
    if locked:
        assert loaded.GetMinimumMaximum() == sdf.GetMinimumMaximum()


@pytest.mark.parametrize("compress", [False, True])
def test_occupancy_map_file_equals_the_oracle_and_round_trips(tmp_path, compress):
    rng = np.random.default_rng(11)
    shape = (6, 3, 5)
    sizes = grids.VoxelGridSizes.FromVoxelCounts(0.5, shape)
    cells = rng.choice(np.array([0.0, 0.5, 1.0], dtype=np.float32), size=shape)
    occupancy_map = grids.OccupancyMap(origin_transform(rng), "world", sizes,
                                       default_occupancy=0.5, data=cells)
    path = tmp_path / "map.cmg"
    grid_files.SaveOccupancyMapToFile(occupancy_map, path, compress)
    content = path.read_bytes()
    assert content[:4] == (b"CMGZ" if compress else b"CMGR")
    payload = zlib.decompress(content[4:]) if compress else content[4:]
    want = gfo.serialize_grid(cells, 0.5, occupancy_map.OriginTransform().T.reshape(-1), 0.5, 0.5,
                              "world", locked=None)
    assert payload == want
    loaded = grid_files.LoadOccupancyMapFromFile(path)
    assert np.array_equal(loaded.GetImmutableRawData(), cells)
    assert loaded.Frame() == "world" and loaded.VoxelXSize() == 0.5
    assert np.array_equal(loaded.OriginTransform(), occupancy_map.OriginTransform())
    assert loaded._default_occupancy == 0.5 and loaded._oob_occupancy == 0.5


def test_errors_are_the_references(tmp_path):
    with pytest.raises(ValueError, match="File does not exist"):
        grid_files.LoadSignedDistanceFieldFromFile(tmp_path / "missing.sdf")
    small = tmp_path / "small.sdf"
    small.write_bytes(b"SD")
    with pytest.raises(ValueError, match="File is too small"):
        grid_files.LoadSignedDistanceFieldFromFile(small)
    with pytest.raises(ValueError, match="File is too small"):
        gfo.load_from_file(small, gfo.SDF_MAGIC)
    wrong = tmp_path / "wrong.sdf"
    wrong.write_bytes(b"CMGR" + b"\0" * 64)
    with pytest.raises(ValueError, match=r"File has invalid header \[CMGR\]"):
        grid_files.LoadSignedDistanceFieldFromFile(wrong)
    with pytest.raises(ValueError, match=r"File has invalid header \[CMGR\]"):
        gfo.load_from_file(wrong, gfo.SDF_MAGIC)
    # a truncated payload is refused, not read past its end
    rng = np.random.default_rng(3)
    sdf = make_sdf(rng, np.float32)
    path = tmp_path / "field.sdf"
    grid_files.SaveSignedDistanceFieldToFile(sdf, path, False)
    cut = tmp_path / "cut.sdf"
    cut.write_bytes(path.read_bytes()[:-9])
    with pytest.raises(ValueError, match="Not enough room"):
        grid_files.LoadSignedDistanceFieldFromFile(cut)
    damaged = tmp_path / "damaged.sdf"
    damaged.write_bytes(b"SDFZ" + b"not a zlib stream")
    with pytest.raises(ValueError, match="zlib"):
        grid_files.LoadSignedDistanceFieldFromFile(damaged)
    # a float64 file read as float32 does not parse (the cell count no longer matches)
    double_path = tmp_path / "double.sdf"
    grid_files.SaveSignedDistanceFieldToFile(make_sdf(rng, np.float64), double_path, False)
    with pytest.raises(ValueError):
        grid_files.LoadSignedDistanceFieldFromFile(double_path, np.float32)


def test_members_named_like_the_references(tmp_path):
    # SignedDistanceField<T>::SaveToFile(sdf, path, compress) / ::LoadFromFile(path) and the same
    # pair on OccupancyMap, as static members of the mirror classes
    rng = np.random.default_rng(13)
    sdf = make_sdf(rng, np.float32)
    path = tmp_path / "member.sdf"
    grids.SignedDistanceField.SaveToFile(sdf, path, True)
    loaded = grids.SignedDistanceField.LoadFromFile(path)
    assert np.array_equal(loaded.GetImmutableRawData(), sdf.GetImmutableRawData())
    # without a dtype the scalar type is found from the file
    for dtype in (np.float32, np.float64):
        typed = tmp_path / f"typed_{np.dtype(dtype).name}.sdf"
        original = make_sdf(rng, dtype, shape=(3, 5, 2))
        grids.SignedDistanceField.SaveToFile(original, typed, False)
        found = grids.SignedDistanceField.LoadFromFile(typed)
        assert found.GetImmutableRawData().dtype == np.dtype(dtype)
        assert np.array_equal(found.GetImmutableRawData(), original.GetImmutableRawData())
    with pytest.raises(ValueError, match="File does not exist"):
        grids.SignedDistanceField.LoadFromFile(tmp_path / "nothing.sdf")
    sizes = grids.VoxelGridSizes.FromVoxelCounts(0.5, (2, 3, 4))
    occupancy_map = grids.OccupancyMap(np.eye(4), "m", sizes, default_occupancy=0.5)
    grids.OccupancyMap.SaveToFile(occupancy_map, tmp_path / "member.cmg", False)
    again = grids.OccupancyMap.LoadFromFile(tmp_path / "member.cmg")
    assert np.array_equal(again.GetImmutableRawData(), occupancy_map.GetImmutableRawData())


def test_empty_frame_and_single_cell(tmp_path):
    sizes = grids.VoxelGridSizes.FromVoxelCounts(1.0, (1, 1, 1))
    sdf = grids.SignedDistanceField(np.eye(4), "", sizes, np.array([[[2.5]]], dtype=np.float32), 0.0)
    path = tmp_path / "one.sdf"
    grid_files.SaveSignedDistanceFieldToFile(sdf, path, True)
    loaded = grid_files.LoadSignedDistanceFieldFromFile(path)
    assert loaded.Frame() == "" and not loaded.IsLocked()
    assert loaded.GetImmutableRawData().tolist() == [[[2.5]]] and loaded._oob_value == 0.0


# ---- committed golden files written by the reference's own members --------------------------------
def test_golden_files_written_by_the_reference(tmp_path):
    # tests/golden/grid_files/*: written by SignedDistanceField<T>::SaveToFile and
    # OccupancyMap::SaveToFile of the reference build (tests/golden/make_grid_file_goldens.py).
    # The product must write the same bytes for the same inputs, and read them back.
    import json
    from pathlib import Path
    golden = Path(__file__).resolve().parent / "golden" / "grid_files"
    index = json.loads((golden / "index.json").read_text())
    shape = tuple(index["shape"])
    origin = np.array(index["origin_row_major"]).reshape(4, 4)
    values = (np.arange(np.prod(shape), dtype=np.float64).reshape(shape) - 7.5) * 0.125
    cells = np.array([0.0, 0.5, 1.0, 0.25], dtype=np.float32)[np.arange(np.prod(shape)) % 4]
    cells = cells.reshape(shape)
    sizes = grids.VoxelGridSizes.FromVoxelCounts(index["resolution"], shape)
    assert len(index["files"]) == 10
    for name, what in index["files"].items():
        want = (golden / name).read_bytes()
        out = tmp_path / name
        if what["kind"] == "sdf":
            dtype = np.float32 if what["dtype"] == "f32" else np.float64
            sdf = grids.SignedDistanceField(origin, index["frame"], sizes, values.astype(dtype),
                                            index["oob_value"])
            if what["locked"]:
                sdf.Lock()
            grid_files.SaveSignedDistanceFieldToFile(sdf, out, what["compressed"])
            assert out.read_bytes() == want, name
            loaded = grid_files.LoadSignedDistanceFieldFromFile(golden / name, dtype)
            assert np.array_equal(loaded.GetImmutableRawData(), values.astype(dtype))
            assert loaded.IsLocked() == what["locked"] and loaded.Frame() == index["frame"]
            assert loaded._oob_value == index["oob_value"]
            assert np.array_equal(loaded.OriginTransform(), origin)
            # the restated oracle agrees with the reference's bytes too
            payload = gfo.load_from_file(golden / name, gfo.SDF_MAGIC)
            assert payload == gfo.serialize_grid(values.astype(dtype), index["resolution"],
                                                 origin.T.reshape(-1), index["oob_value"],
                                                 index["oob_value"], index["frame"],
                                                 locked=what["locked"])
        else:
            occupancy_map = grids.OccupancyMap(origin, index["frame"], sizes,
                                               default_occupancy=what["default_occupancy"],
                                               data=cells)
            occupancy_map._oob_occupancy = what["oob_occupancy"]
            grid_files.SaveOccupancyMapToFile(occupancy_map, out, what["compressed"])
            assert out.read_bytes() == want, name
            loaded = grid_files.LoadOccupancyMapFromFile(golden / name)
            assert np.array_equal(loaded.GetImmutableRawData(), cells)
            assert loaded._default_occupancy == what["default_occupancy"]
            assert loaded._oob_occupancy == what["oob_occupancy"]


# ---- against the reference's own members (oracle/_ref, built where /root/reference exists) ----
def reference_save(reference_library, sdf, path, compress):
    data = np.ascontiguousarray(sdf.GetImmutableRawData())
    origin = np.ascontiguousarray(sdf.OriginTransform().T.reshape(-1))
    message = ctypes.create_string_buffer(256)
    code = reference_library.vgt_ref_sdf_save_to_file(
        ctypes.c_int(data.dtype.itemsize), data.ctypes.data_as(ctypes.c_void_p),
        *(ctypes.c_int64(v) for v in data.shape), ctypes.c_double(sdf.Resolution()),
        origin.ctypes.data_as(ctypes.c_void_p), sdf.Frame().encode(), ctypes.c_int(sdf.IsLocked()),
        ctypes.c_double(sdf._oob_value), str(path).encode(), ctypes.c_int(compress), message,
        ctypes.c_int64(256))
    assert code == 0, message.value


def reference_load(reference_library, path, dtype, capacity=4096):
    values = np.zeros(capacity, dtype=dtype)
    dims = (ctypes.c_int64 * 3)()
    resolution = ctypes.c_double()
    origin = np.zeros(16)
    frame = ctypes.create_string_buffer(128)
    locked = ctypes.c_int()
    default_and_oob = (ctypes.c_double * 2)()
    min_max = (ctypes.c_double * 2)()
    message = ctypes.create_string_buffer(256)
    code = reference_library.vgt_ref_sdf_load_from_file(
        ctypes.c_int(np.dtype(dtype).itemsize), str(path).encode(),
        values.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(capacity), dims,
        ctypes.byref(resolution), origin.ctypes.data_as(ctypes.c_void_p), frame,
        ctypes.c_int64(128), ctypes.byref(locked), default_and_oob, min_max, message,
        ctypes.c_int64(256))
    return code, message.value.decode(), {
        "cells": values[:dims[0] * dims[1] * dims[2]].reshape(tuple(dims)),
        "resolution": resolution.value, "origin": origin.reshape(4, 4).T, "frame": frame.value.decode(),
        "locked": bool(locked.value), "oob": default_and_oob[1], "min_max": tuple(min_max)}


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("compress", [False, True])
@pytest.mark.parametrize("locked", [False, True])
def test_files_equal_the_references_own_members(tmp_path, reference_library, dtype, compress,
                                                locked):
    rng = np.random.default_rng(19)
    sdf = make_sdf(rng, dtype, shape=(4, 6, 3), locked=locked, oob=7.5)
    ours, theirs = tmp_path / "ours.sdf", tmp_path / "theirs.sdf"
    grid_files.SaveSignedDistanceFieldToFile(sdf, ours, compress)
    reference_save(reference_library, sdf, theirs, compress)
    # byte for byte: magic, compression, grid bytes, frame, locked flag
    assert ours.read_bytes() == theirs.read_bytes()
    # the reference's LoadFromFile reads our file ...
    code, message, got = reference_load(reference_library, ours, dtype)
    assert code == 0, message
    assert np.array_equal(got["cells"], sdf.GetImmutableRawData())
    assert got["frame"] == "some_frame" and got["locked"] == locked and got["oob"] == 7.5
    assert got["resolution"] == 0.25 and np.array_equal(got["origin"], sdf.OriginTransform())
    if locked:
        assert got["min_max"] == tuple(float(v) for v in sdf.GetMinimumMaximum())
    # ... and we read the reference's
    loaded = grid_files.LoadSignedDistanceFieldFromFile(theirs, dtype)
    assert np.array_equal(loaded.GetImmutableRawData(), sdf.GetImmutableRawData())
    assert loaded.IsLocked() == locked and loaded.Frame() == "some_frame"


def test_reference_errors_match(tmp_path, reference_library):
    code, message, _ = reference_load(reference_library, tmp_path / "missing.sdf", np.float32)
    assert code == 1 and message == "File does not exist"
    small = tmp_path / "small.sdf"
    small.write_bytes(b"SD")
    code, message, _ = reference_load(reference_library, small, np.float32)
    assert code == 1 and message == "File is too small"
    wrong = tmp_path / "wrong.sdf"
    wrong.write_bytes(b"CMGR" + b"\0" * 64)
    code, message, _ = reference_load(reference_library, wrong, np.float32)
    assert code == 1 and message == "File has invalid header [CMGR]"


@pytest.fixture(scope="module")
def reference_maps_library():
    """oracle/_ref/libvgt_ref_maps.so: the reference's own OccupancyMap with its file members."""
    path = reference_oracle._PATH.with_name("libvgt_ref_maps.so")
    if not path.exists():
        pytest.skip("oracle/_ref/libvgt_ref_maps.so is not built (needs /root/reference)")
    return ctypes.CDLL(str(path))


@pytest.mark.parametrize("compress", [False, True])
def test_map_files_equal_the_references_own_members(tmp_path, reference_maps_library, compress):
    rng = np.random.default_rng(23)
    shape = (5, 7, 3)
    sizes = grids.VoxelGridSizes.FromVoxelCounts(0.125, shape)
    cells = rng.choice(np.array([0.0, 0.25, 0.5, 1.0], dtype=np.float32), size=shape)
    occupancy_map = grids.OccupancyMap(origin_transform(rng), "map_frame", sizes,
                                       default_occupancy=0.5, data=cells)
    ours, theirs = tmp_path / "ours.cmg", tmp_path / "theirs.cmg"
    grid_files.SaveOccupancyMapToFile(occupancy_map, ours, compress)
    origin = np.ascontiguousarray(occupancy_map.OriginTransform().T.reshape(-1))
    message = ctypes.create_string_buffer(256)
    code = reference_maps_library.vgt_ref_map_save_to_file(
        cells.ctypes.data_as(ctypes.c_void_p), *(ctypes.c_int64(v) for v in shape),
        ctypes.c_double(0.125), origin.ctypes.data_as(ctypes.c_void_p), b"map_frame",
        ctypes.c_float(0.5), ctypes.c_float(0.5), str(theirs).encode(), ctypes.c_int(compress),
        message, ctypes.c_int64(256))
    assert code == 0, message.value
    assert ours.read_bytes() == theirs.read_bytes()
    # the reference's LoadFromFile on our file
    values = np.zeros(cells.size, dtype=np.float32)
    dims = (ctypes.c_int64 * 3)()
    resolution = ctypes.c_double()
    got_origin = np.zeros(16)
    frame = ctypes.create_string_buffer(64)
    default_and_oob = (ctypes.c_float * 2)()
    code = reference_maps_library.vgt_ref_map_load_from_file(
        str(ours).encode(), values.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(values.size),
        dims, ctypes.byref(resolution), got_origin.ctypes.data_as(ctypes.c_void_p), frame,
        ctypes.c_int64(64), default_and_oob, message, ctypes.c_int64(256))
    assert code == 0, message.value
    assert tuple(dims) == shape and resolution.value == 0.125
    assert frame.value == b"map_frame" and tuple(default_and_oob) == (0.5, 0.5)
    assert np.array_equal(values.reshape(shape), cells)
    assert np.array_equal(got_origin.reshape(4, 4).T, occupancy_map.OriginTransform())
    # ours on the reference's file
    loaded = grid_files.LoadOccupancyMapFromFile(theirs)
    assert np.array_equal(loaded.GetImmutableRawData(), cells) and loaded.Frame() == "map_frame"
    # and the reference's errors
    code = reference_maps_library.vgt_ref_map_load_from_file(
        str(tmp_path / "missing.cmg").encode(), values.ctypes.data_as(ctypes.c_void_p),
        ctypes.c_int64(values.size), dims, ctypes.byref(resolution),
        got_origin.ctypes.data_as(ctypes.c_void_p), frame, ctypes.c_int64(64), default_and_oob,
        message, ctypes.c_int64(256))
    assert code == 1 and message.value == b"File does not exist"
    with pytest.raises(ValueError, match="File does not exist"):
        grid_files.LoadOccupancyMapFromFile(tmp_path / "missing.cmg")


@pytest.mark.gpu
def test_device_resident_sdf_saved_straight_to_a_file(tmp_path):
    import torch
    rng = np.random.default_rng(5)
    data = rng.normal(size=(9, 8, 16)).astype(np.float32)
    d_sdf = torch.from_numpy(data).cuda()
    path = tmp_path / "device.sdf"
    grid_files.SaveDeviceSignedDistanceFieldToFile(d_sdf, 0.1, np.eye(4), "dev", path, True)
    loaded = grid_files.LoadSignedDistanceFieldFromFile(path)
    assert np.array_equal(loaded.GetImmutableRawData(), data)
    assert loaded.IsLocked() and loaded.Frame() == "dev" and loaded.Resolution() == 0.1
