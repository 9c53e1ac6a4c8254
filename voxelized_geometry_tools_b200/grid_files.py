"""The reference's grid files through the C-ABI (vgt_b200_grid_file_*, csrc/grid_files.cu).

Mirrors ``SignedDistanceField<T>::SaveToFile / LoadFromFile``
(include/voxelized_geometry_tools/signed_distance_field.hpp:643-722) and
``OccupancyMap::SaveToFile / LoadFromFile`` (src/voxelized_geometry_tools/occupancy_map.cpp:116-193):
same names, same argument meaning, ``ValueError`` where the reference throws
``std::invalid_argument`` ("File does not exist", "File is too small", "File has invalid header").
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _capi
from .grids import OccupancyMap, SignedDistanceField, VoxelGridSizes


def _info(shape, voxel_size, origin_transform, default_value, oob_value, locked):
    info = _capi.GridFileInfo()
    info.nx, info.ny, info.nz = (int(v) for v in shape)
    for i in range(3):
        info.voxel_size[i] = float(voxel_size)
    column_major = np.asarray(origin_transform, dtype=np.float64).reshape(4, 4).T.reshape(-1)
    for i in range(16):
        info.origin_transform[i] = float(column_major[i])
    info.default_value = float(default_value)
    info.oob_value = float(oob_value)
    info.initialized = 1
    info.locked = 1 if locked else 0
    return info


def _save(path, kind, compress, data, info, frame):
    cells = np.ascontiguousarray(data)
    _capi.check(_capi.library().vgt_b200_grid_file_save(
        os.fsencode(path), kind, 1 if compress else 0, cells.ctypes.data_as(ctypes.c_void_p),
        ctypes.byref(info), frame.encode("utf-8")))


def _load(path, kind, dtype):
    library = _capi.library()
    info = _capi.GridFileInfo()
    _capi.check(library.vgt_b200_grid_file_probe(os.fsencode(path), kind, ctypes.byref(info), None, 0))
    frame = ctypes.create_string_buffer(int(info.frame_length) + 1)
    cells = np.empty((int(info.nx), int(info.ny), int(info.nz)), dtype=dtype)
    _capi.check(library.vgt_b200_grid_file_load(
        os.fsencode(path), kind, cells.ctypes.data_as(ctypes.c_void_p), cells.size,
        ctypes.byref(info), frame, len(frame)))
    origin = np.array(list(info.origin_transform), dtype=np.float64).reshape(4, 4).T.copy()
    sizes = VoxelGridSizes.FromVoxelCounts(float(info.voxel_size[0]),
                                           (int(info.nx), int(info.ny), int(info.nz)))
    return cells, origin, sizes, frame.value.decode("utf-8"), info


def _sdf_kind(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return _capi.GRID_FILE_SDF_F32
    if dtype == np.float64:
        return _capi.GRID_FILE_SDF_F64
    raise ValueError("SignedDistanceField files hold float32 or float64 cells")


def SaveSignedDistanceFieldToFile(sdf: SignedDistanceField, filepath, compress: bool) -> None:
    """SignedDistanceField<T>::SaveToFile (signed_distance_field.hpp:643-668)."""
    data = sdf.GetImmutableRawData()
    info = _info(data.shape, sdf.Resolution(), sdf.OriginTransform(), sdf._oob_value,
                 sdf._oob_value, sdf.IsLocked())
    _save(filepath, _sdf_kind(data.dtype), compress, data, info, sdf.Frame())


def LoadSignedDistanceFieldFromFile(filepath, dtype=None) -> SignedDistanceField:
    """SignedDistanceField<T>::LoadFromFile (signed_distance_field.hpp:670-722); a file saved
    locked comes back locked, with its extrema recomputed like Lock() does (:579-592).
    dtype plays the part of the reference's template argument; None = whichever of float32 /
    float64 the file parses as (the format does not name its scalar type, but the cell count
    and the voxel counts only agree for the right one)."""
    if dtype is None:
        try:
            return LoadSignedDistanceFieldFromFile(filepath, np.float32)
        except ValueError as first_error:
            if str(first_error).startswith("File"):   # missing / too small / wrong magic
                raise
            return LoadSignedDistanceFieldFromFile(filepath, np.float64)
    cells, origin, sizes, frame, info = _load(filepath, _sdf_kind(dtype), np.dtype(dtype))
    sdf = SignedDistanceField(origin, frame, sizes, cells, info.oob_value)
    if info.locked:
        sdf.Lock()
    return sdf


def SaveOccupancyMapToFile(occupancy_map: OccupancyMap, filepath, compress: bool) -> None:
    """OccupancyMap::SaveToFile (occupancy_map.cpp:116-141)."""
    data = occupancy_map.GetImmutableRawData()
    info = _info(data.shape, occupancy_map.VoxelXSize(), occupancy_map.OriginTransform(),
                 occupancy_map._default_occupancy, occupancy_map._oob_occupancy, False)
    _save(filepath, _capi.GRID_FILE_OCCUPANCY, compress, data, info, occupancy_map.Frame())


def LoadOccupancyMapFromFile(filepath) -> OccupancyMap:
    """OccupancyMap::LoadFromFile (occupancy_map.cpp:143-193)."""
    cells, origin, sizes, frame, info = _load(filepath, _capi.GRID_FILE_OCCUPANCY, np.float32)
    loaded = OccupancyMap(origin, frame, sizes, default_occupancy=float(info.default_value),
                          data=cells)
    loaded._oob_occupancy = float(info.oob_value)
    return loaded


def SaveDeviceSignedDistanceFieldToFile(d_sdf, resolution: float, origin_transform, frame: str,
                                        filepath, compress: bool, oob_value=float("inf"),
                                        locked: bool = True) -> None:
    """A float32 / float64 SDF that lives on the device (a torch CUDA tensor [x, y, z]) written
    as the file SignedDistanceField<T>::SaveToFile would write for it."""
    import torch
    if not (isinstance(d_sdf, torch.Tensor) and d_sdf.is_cuda and d_sdf.is_contiguous()
            and d_sdf.dim() == 3):
        raise ValueError("expected a contiguous 3-D CUDA tensor")
    kind = _sdf_kind({torch.float32: np.float32, torch.float64: np.float64}[d_sdf.dtype])
    info = _info(tuple(d_sdf.shape), resolution, origin_transform, oob_value, oob_value, locked)
    stream = torch.cuda.current_stream(d_sdf.device).cuda_stream
    _capi.check(_capi.library().vgt_b200_grid_file_save_dev(
        os.fsencode(filepath), kind, 1 if compress else 0, ctypes.c_void_p(d_sdf.data_ptr()),
        ctypes.byref(info), frame.encode("utf-8"), d_sdf.device.index or 0,
        ctypes.c_void_p(stream)))
