"""Host-side mirror of the reference's mesh rasterizer interface
(include/voxelized_geometry_tools/mesh_rasterizer.hpp:19-91,
src/voxelized_geometry_tools/mesh_rasterizer.cpp): same function names, argument meaning and
error behaviour, the work done by csrc/mesh_rasterizer.cu through the C-ABI.

    RasterizeMesh(vertices, triangles, occupancy_map, enforce_occupancy_map_contains_mesh)
    RasterizeTriangle(vertices, triangles, triangle_index, occupancy_map, enforce...)
    RasterizeMeshIntoOccupancyMap(vertices, triangles, resolution)
    RasterizeMeshIntoOccupancyComponentMap(vertices, triangles, resolution)

vertices: float64 [n, 3]; triangles: int32 [m, 3]. Errors: ValueError where the reference throws
std::invalid_argument, RuntimeError for "Triangle is not contained by occupancy map",
IndexError where vertices.at() / triangles.at() throw std::out_of_range. The `parallelism`
argument of the reference is accepted and ignored (the device decides).
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _capi
from .grids import OccupancyComponentMap, OccupancyMap, VoxelGridSizes, inverse_rigid


def _mesh_arrays(vertices, triangles):
    vertices = np.ascontiguousarray(vertices, dtype=np.float64)
    triangles = np.ascontiguousarray(triangles, dtype=np.int32)
    if vertices.ndim != 2 or vertices.shape[1] != 3:
        raise ValueError("vertices must be an [n, 3] array")
    if triangles.ndim != 2 or triangles.shape[1] != 3:
        raise ValueError("triangles must be an [m, 3] array")
    return vertices, triangles


def _column_major(transform) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(transform, dtype=np.float64).reshape(4, 4).T).reshape(16)


def RasterizeMesh(vertices, triangles, occupancy_map, enforce_occupancy_map_contains_mesh: bool,
                  parallelism=None, device: int = 0) -> None:
    """mesh_rasterizer.cpp:205-230 (OccupancyMap and OccupancyComponentMap overloads :299-321)."""
    vertices, triangles = _mesh_arrays(vertices, triangles)
    cells = occupancy_map.GetMutableRawData()
    if not cells.flags.c_contiguous:
        raise ValueError("occupancy_map storage must be contiguous")
    _capi.require_device(device)
    x_wg = _column_major(occupancy_map.OriginTransform())
    x_gw = _column_major(inverse_rigid(occupancy_map.OriginTransform()))
    code = _capi.library().vgt_b200_rasterize_mesh_f64(
        vertices.ctypes.data, len(vertices), triangles.ctypes.data, len(triangles),
        cells.ctypes.data, cells.dtype.itemsize, *cells.shape,
        float(occupancy_map.ControlSizes().voxel_size),
        x_wg.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        x_gw.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        int(bool(enforce_occupancy_map_contains_mesh)), device)
    _capi.check(code)


def RasterizeTriangle(vertices, triangles, triangle_index: int, occupancy_map,
                      enforce_occupancy_map_contains_triangle: bool, device: int = 0) -> None:
    """mesh_rasterizer.cpp:283-297: one triangle of the list."""
    vertices, triangles = _mesh_arrays(vertices, triangles)
    if not 0 <= int(triangle_index) < len(triangles):
        raise IndexError("triangle_index out of range")      # triangles.at(triangle_index)
    RasterizeMesh(vertices, triangles[int(triangle_index):int(triangle_index) + 1], occupancy_map,
                  enforce_occupancy_map_contains_triangle, device=device)


def _map_around(vertices, resolution: float):
    """The grid RasterizeMeshInto...MapImpl sizes around the mesh (mesh_rasterizer.cpp:239-272):
    the mesh's bounding box plus one voxel on every side."""
    if not resolution > 0.0:
        raise ValueError("resolution must be greater than zero")
    lower = np.full(3, math.inf)
    upper = np.full(3, -math.inf)
    if len(vertices):
        lower = np.minimum(lower, vertices.min(axis=0))
        upper = np.maximum(upper, vertices.max(axis=0))
    object_size = upper - lower
    buffer_size = resolution * 2.0
    sizes = VoxelGridSizes.FromGridSizes(resolution, [float(s + buffer_size) for s in object_size])
    origin = np.eye(4)
    origin[:3, 3] = lower - resolution
    return origin, sizes


def RasterizeMeshIntoOccupancyMap(vertices, triangles, resolution: float, parallelism=None,
                                  device: int = 0) -> OccupancyMap:
    """mesh_rasterizer.cpp:232-279, :323-331."""
    vertices, triangles = _mesh_arrays(vertices, triangles)
    origin, sizes = _map_around(vertices, float(resolution))
    occupancy_map = OccupancyMap(origin, "mesh", sizes, 0.0)
    RasterizeMesh(vertices, triangles, occupancy_map, True, device=device)
    return occupancy_map


def RasterizeMeshIntoOccupancyComponentMap(vertices, triangles, resolution: float,
                                           parallelism=None,
                                           device: int = 0) -> OccupancyComponentMap:
    """mesh_rasterizer.cpp:333-342."""
    vertices, triangles = _mesh_arrays(vertices, triangles)
    origin, sizes = _map_around(vertices, float(resolution))
    occupancy_map = OccupancyComponentMap(origin, "mesh", sizes)
    RasterizeMesh(vertices, triangles, occupancy_map, True, device=device)
    return occupancy_map
