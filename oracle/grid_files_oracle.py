"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy + zlib) of the reference's grid files.

Only tests/ may import this module; the product (csrc/grid_files.cu behind
vgt_b200_grid_file_*) never does.

Follows, statement by statement,
  * SignedDistanceField<T>::SaveToFile / LoadFromFile
    (include/voxelized_geometry_tools/signed_distance_field.hpp:643-722): four-byte magic "SDFZ"
    (one zlib stream of the serialized field) or "SDFR" (the serialized field as it is);
    "File does not exist" / "File is too small" / "File has invalid header [....]";
  * its derived members (signed_distance_field.hpp:551-596): the frame name as a string, then
    the locked flag as one byte, appended AFTER the grid's own bytes; a field saved locked is
    locked again on load;
  * OccupancyMap::SaveToFile / LoadFromFile (src/voxelized_geometry_tools/occupancy_map.cpp:
    116-193, "CMGZ" / "CMGR"), derived member = the frame name only (:56-84), one float per cell
    (:23-46).

PARITY UNPINNED below that: the grid's own bytes and the primitive encodings belong to
common_robotics_utilities (VoxelGridBase::SerializeSelf, serialization.hpp, zlib_helpers.hpp),
which is neither in the reference tree nor pinned by it (package.xml.ros2:12). The layout is
restated as published by that library: raw little-endian items; strings and vectors as a uint64
count followed by the items; an Isometry3d as its 4x4 matrix, column-major; the grid as
initialized flag, origin transform, inverse origin transform, cells, voxel sizes, voxel counts,
default value, out-of-bounds value. The reference's own tests never touch these members; the
golden files under tests/golden/grid_files/ were written by the reference's members compiled
over the same restated third-party layer. tests/test_grid_files.py checks the product against this module and against
the reference's own members - SignedDistanceField<T> and OccupancyMap, headers and sources
unmodified - compiled over oracle/ref_shim (which restates the same third-party layer).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SDF_MAGIC = (b"SDFR", b"SDFZ")
MAP_MAGIC = (b"CMGR", b"CMGZ")


def inverse_rigid_column_major(origin_column_major):
    m = np.asarray(origin_column_major, dtype=np.float64).reshape(4, 4).T  # m[r, c]
    out = np.zeros((4, 4))
    out[:3, :3] = m[:3, :3].T
    for r in range(3):
        out[r, 3] = -((out[r, 0] * m[0, 3] + out[r, 1] * m[1, 3]) + out[r, 2] * m[2, 3])
    out[3, 3] = 1.0
    return out.T.reshape(-1)


def serialize_grid(cells: np.ndarray, voxel_size: float, origin_column_major, default_value,
                   oob_value, frame: str, locked=None) -> bytes:
    """The serialized grid: base form, then the derived members (locked is None: an occupancy
    map, which has no such member)."""
    cells = np.ascontiguousarray(cells)
    scalar = {np.dtype(np.float32): "<f", np.dtype(np.float64): "<d"}[cells.dtype]
    origin = np.asarray(origin_column_major, dtype="<f8").reshape(16)
    parts = [struct.pack("<B", 1), origin.tobytes(),
             inverse_rigid_column_major(origin).astype("<f8").tobytes(),
             struct.pack("<Q", cells.size), cells.astype(cells.dtype.newbyteorder("<")).tobytes(),
             struct.pack("<ddd", voxel_size, voxel_size, voxel_size),
             struct.pack("<qqq", *cells.shape),
             struct.pack(scalar, default_value), struct.pack(scalar, oob_value)]
    name = frame.encode("utf-8")
    parts += [struct.pack("<Q", len(name)), name]
    if locked is not None:
        parts.append(struct.pack("<B", 1 if locked else 0))
    return b"".join(parts)


def deserialize_grid(payload: bytes, dtype, has_locked: bool):
    dtype = np.dtype(dtype)
    at = 0

    def take(fmt):
        nonlocal at
        size = struct.calcsize(fmt)
        if at + size > len(payload):
            raise ValueError("Not enough room in the provided buffer")
        values = struct.unpack_from(fmt, payload, at)
        at += size
        return values

    (initialized,) = take("<B")
    origin = np.array(take("<16d"))
    inverse = np.array(take("<16d"))
    (count,) = take("<Q")
    size = count * dtype.itemsize
    if at + size > len(payload):
        raise ValueError("Not enough room in the provided buffer")
    flat = np.frombuffer(payload, dtype=dtype.newbyteorder("<"), count=count, offset=at)
    at += size
    voxel_sizes = take("<ddd")
    counts = take("<qqq")
    scalar = "<f" if dtype.itemsize == 4 else "<d"
    (default_value,) = take(scalar)
    (oob_value,) = take(scalar)
    (frame_length,) = take("<Q")
    if at + frame_length > len(payload):
        raise ValueError("Not enough room in the provided buffer")
    frame = payload[at:at + frame_length].decode("utf-8")
    at += frame_length
    locked = None
    if has_locked:
        (locked,) = take("<B")
    if counts[0] * counts[1] * counts[2] != count:
        raise ValueError("serialized grid holds the wrong number of cells")
    return {"initialized": bool(initialized), "origin": origin, "inverse_origin": inverse,
            "cells": flat.reshape(counts).astype(dtype), "voxel_sizes": voxel_sizes,
            "default_value": default_value, "oob_value": oob_value, "frame": frame,
            "locked": None if locked is None else bool(locked), "bytes_read": at}


def save_to_file(path, magics, payload: bytes, compress: bool) -> None:
    with open(path, "wb") as output_file:
        if compress:
            output_file.write(magics[1])
            output_file.write(zlib.compress(payload))      # one stream, default level
        else:
            output_file.write(magics[0])
            output_file.write(payload)


def load_from_file(path, magics) -> bytes:
    try:
        with open(path, "rb") as input_file:
            content = input_file.read()
    except FileNotFoundError:
        raise ValueError("File does not exist") from None
    if len(content) < 4:
        raise ValueError("File is too small")
    header, body = content[:4], content[4:]
    if header == magics[1]:
        return zlib.decompress(body)
    if header == magics[0]:
        return body
    raise ValueError("File has invalid header [" + header.decode("latin-1") + "]")
