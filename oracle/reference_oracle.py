"""ctypes front-end for oracle/_ref/libvgt_ref.so (TEST INFRASTRUCTURE ONLY).

That library is the REFERENCE'S OWN signed_distance_field_generation.{hpp,cpp}, compiled
unmodified from /root/reference over the stand-in headers in oracle/ref_shim/ (``make -C oracle
ref``). It is used (a) to pin the restated oracle (tests/test_oracle_vs_reference.py) and (b) as
the CPU baseline of kind "reference" in bench.py. It is built in the dev container and travels
to the GPU box prebuilt (oracle/_ref/ is git-ignored but not gpurun-ignored).
"""
from __future__ import annotations

import ctypes
from pathlib import Path

import numpy as np

_PATH = Path(__file__).resolve().parent / "_ref" / "libvgt_ref.so"
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64
_int = ctypes.c_int
_lib = None


def available() -> bool:
    return _PATH.exists()


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        handle = ctypes.CDLL(str(_PATH))
        handle.vgt_ref_sdf_f32.argtypes = [_f32p, _i64, _i64, _i64, ctypes.c_double, _int, _int,
                                           _int, _f32p, _f32p]
        handle.vgt_ref_sdf_f64.argtypes = [_f32p, _i64, _i64, _i64, ctypes.c_double, _int, _int,
                                           _int, _f64p, _f64p]
        handle.vgt_ref_transform_inplace_f64.argtypes = [_f64p, _i64, _i64, _i64, _int]
        _lib = handle
    return _lib


def sdf(occupancy, resolution: float, unknown_is_filled: bool = True,
        add_virtual_border: bool = False, threads: int = 0, dtype=np.float32):
    occ = np.ascontiguousarray(occupancy, dtype=np.float32)
    out = np.empty(occ.shape, dtype=dtype)
    min_max = np.zeros(2, dtype=dtype)
    if np.dtype(dtype) == np.float32:
        code = lib().vgt_ref_sdf_f32(
            occ.ctypes.data_as(_f32p), *occ.shape, float(resolution), int(unknown_is_filled),
            int(add_virtual_border), threads, out.ctypes.data_as(_f32p),
            min_max.ctypes.data_as(_f32p))
    else:
        code = lib().vgt_ref_sdf_f64(
            occ.ctypes.data_as(_f32p), *occ.shape, float(resolution), int(unknown_is_filled),
            int(add_virtual_border), threads, out.ctypes.data_as(_f64p),
            min_max.ctypes.data_as(_f64p))
    if code != 0:
        raise RuntimeError("reference SDF generation failed")
    return out, (min_max[0], min_max[1])


# The reference's own OccupancyMap (occupancy_map.hpp / .cpp unmodified) lives in a library of
# its own, see oracle/ref_shim/ref_map_files_entry.cpp.
_MAPS_PATH = _PATH.with_name("libvgt_ref_maps.so")
_maps_lib = None


def maps_available() -> bool:
    return _MAPS_PATH.exists()


def occupancy_map_sdf(occupancy, resolution: float, unknown_is_filled: bool = True,
                      add_virtual_border: bool = False, threads: int = 0, dtype=np.float32):
    """OccupancyMap::ExtractSignedDistanceField<T> (occupancy_map.hpp:174-216) - the public entry
    this backend drops in for - on the reference's own class."""
    global _maps_lib
    if _maps_lib is None:
        _maps_lib = ctypes.CDLL(str(_MAPS_PATH))
    occ = np.ascontiguousarray(occupancy, dtype=np.float32)
    out = np.empty(occ.shape, dtype=dtype)
    min_max = np.zeros(2, dtype=dtype)
    code = _maps_lib.vgt_ref_map_extract_sdf(
        ctypes.c_int(np.dtype(dtype).itemsize), occ.ctypes.data_as(ctypes.c_void_p),
        *(ctypes.c_int64(v) for v in occ.shape), ctypes.c_double(resolution),
        ctypes.c_int(unknown_is_filled), ctypes.c_int(add_virtual_border), ctypes.c_int(threads),
        out.ctypes.data_as(ctypes.c_void_p), min_max.ctypes.data_as(ctypes.c_void_p))
    if code != 0:
        raise RuntimeError("reference OccupancyMap::ExtractSignedDistanceField failed")
    return out, (min_max[0], min_max[1])


def tagged_map_sdf(cells, resolution: float, objects_to_use=(), unknown_is_filled: bool = True,
                   add_virtual_border: bool = False, free_and_named: bool = False,
                   dtype=np.float32):
    """TaggedObjectOccupancyMap::ExtractSignedDistanceField<T>(objects_to_use, parameters)
    (tagged_object_occupancy_map.hpp:199-247) or, free_and_named=True,
    ::ExtractFreeAndNamedObjectsSignedDistanceField<T> (:293-378), on the reference's own class.
    cells: structured array with 8-byte {occupancy float32, object_id uint32} items."""
    global _maps_lib
    if _maps_lib is None:
        _maps_lib = ctypes.CDLL(str(_MAPS_PATH))
    cells = np.ascontiguousarray(cells)
    assert cells.dtype.itemsize == 8, "TaggedObjectOccupancyCell is {float32, uint32}"
    ids = np.ascontiguousarray(np.asarray(list(objects_to_use), dtype=np.uint32))
    out = np.empty(cells.shape, dtype=dtype)
    min_max = np.zeros(2, dtype=dtype)
    code = _maps_lib.vgt_ref_tagged_map_sdf(
        ctypes.c_int(np.dtype(dtype).itemsize), cells.ctypes.data_as(ctypes.c_void_p),
        *(ctypes.c_int64(v) for v in cells.shape), ctypes.c_double(resolution),
        ids.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(ids.size),
        ctypes.c_int(unknown_is_filled), ctypes.c_int(add_virtual_border),
        ctypes.c_int(free_and_named), out.ctypes.data_as(ctypes.c_void_p),
        min_max.ctypes.data_as(ctypes.c_void_p))
    if code != 0:
        raise RuntimeError("reference TaggedObjectOccupancyMap SDF failed")
    return out, (min_max[0], min_max[1])


def component_map_sdf(cells, resolution: float, objects_to_use=(), unknown_is_filled: bool = True,
                      add_virtual_border: bool = False, free_and_named: bool = False,
                      dtype=np.float32):
    """The SDF members of the reference's own OccupancyComponentMap (8-byte cells:
    occupancy_component_map.hpp:270-306) and TaggedObjectOccupancyComponentMap (16-byte cells:
    tagged_object_occupancy_component_map.hpp:360-575), chosen by the cell size."""
    global _maps_lib
    if _maps_lib is None:
        _maps_lib = ctypes.CDLL(str(_MAPS_PATH))
    cells = np.ascontiguousarray(cells)
    if cells.dtype.itemsize == 8:
        kind = 1
        assert not objects_to_use and not free_and_named
    else:
        assert cells.dtype.itemsize == 16
        kind = 3 if free_and_named else 2
    ids = np.ascontiguousarray(np.asarray(list(objects_to_use), dtype=np.uint32))
    out = np.empty(cells.shape, dtype=dtype)
    min_max = np.zeros(2, dtype=dtype)
    code = _maps_lib.vgt_ref_component_map_sdf(
        ctypes.c_int(kind), ctypes.c_int(np.dtype(dtype).itemsize),
        cells.ctypes.data_as(ctypes.c_void_p), *(ctypes.c_int64(v) for v in cells.shape),
        ctypes.c_double(resolution), ids.ctypes.data_as(ctypes.c_void_p),
        ctypes.c_int64(ids.size), ctypes.c_int(unknown_is_filled),
        ctypes.c_int(add_virtual_border), out.ctypes.data_as(ctypes.c_void_p),
        min_max.ctypes.data_as(ctypes.c_void_p))
    if code != 0:
        raise RuntimeError("reference component map SDF failed")
    return out, (min_max[0], min_max[1])


def transform_inplace(field: np.ndarray, threads: int = 0) -> np.ndarray:
    assert field.dtype == np.float64 and field.flags.c_contiguous and field.ndim == 3
    if lib().vgt_ref_transform_inplace_f64(field.ctypes.data_as(_f64p), *field.shape, threads):
        raise RuntimeError("reference transform failed")
    return field


# ------------------------------------------------------------------------------------------------
# The reference's own CPU voxelizer (cpu_pointcloud_voxelization.cpp compiled unmodified; the
# entry points are in oracle/ref_shim/ref_voxelizer_entry.cpp).
# ------------------------------------------------------------------------------------------------
_i32p = ctypes.POINTER(ctypes.c_int32)
_voxelizer_bound = False


def _voxelizer_lib() -> ctypes.CDLL:
    global _voxelizer_bound
    handle = lib()
    if not _voxelizer_bound:
        handle.vgt_ref_voxelize_f64.argtypes = [
            _f32p, _i64, _i64, _i64, ctypes.c_double, ctypes.c_int32,
            ctypes.POINTER(_f64p), ctypes.POINTER(_i64), _f64p, _f64p, ctypes.c_double,
            ctypes.c_int32, ctypes.c_int32, _int, _f32p, _i32p]
        handle.vgt_ref_voxelize_posed_f64.argtypes = [
            _f32p, _i64, _i64, _i64, ctypes.c_double, _f64p, ctypes.c_int32,
            ctypes.POINTER(_f64p), ctypes.POINTER(_i64), _f64p, _f64p, ctypes.c_double,
            ctypes.c_int32, ctypes.c_int32, _int, _f32p]
        handle.vgt_ref_raycast_single_f64.argtypes = [
            _f64p, _f64p, ctypes.c_double, _i64, _i64, _i64, ctypes.c_double, _i32p]
        handle.vgt_ref_isometry_product.argtypes = [_f64p, _f64p, _f64p]
        handle.vgt_ref_isometry_product.restype = None
        handle.vgt_ref_isometry_inverse.argtypes = [_f64p, _f64p]
        handle.vgt_ref_isometry_inverse.restype = None
        _voxelizer_bound = True
    return handle


def voxelizer_available() -> bool:
    if not available():
        return False
    try:
        return hasattr(lib(), "vgt_ref_voxelize_f64")
    except OSError:
        return False


def _cloud_arrays(clouds):
    points = [np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 3) for p, _, _ in clouds]
    pointers = (_f64p * max(1, len(clouds)))(*[p.ctypes.data_as(_f64p) for p in points])
    sizes = (_i64 * max(1, len(clouds)))(*[p.shape[0] for p in points])
    poses = np.ascontiguousarray(
        np.stack([np.asarray(x, dtype=np.float64).reshape(4, 4).T for _, x, _ in clouds])
        if clouds else np.zeros((1, 4, 4)))
    ranges = np.ascontiguousarray([float(r) for _, _, r in clouds] or [0.0], dtype=np.float64)
    return points, pointers, sizes, poses, ranges


def voxelize(static_occupancy, clouds, voxel_size: float, percent_seen_free: float,
             outlier_points_threshold: int, num_cameras_seen_free: int, threads: int = 0):
    """clouds = [(points_xyz, x_gc 4x4, max_range), ...] -> (filtered occupancy,
    counts[cloud, x, y, z, 2]); same contract as oracle.voxelize, computed by the reference's
    RaycastPointCloud / CombineAndFilterGrids."""
    occ = np.ascontiguousarray(static_occupancy, dtype=np.float32)
    out = np.empty_like(occ)
    counts = np.zeros((len(clouds),) + occ.shape + (2,), dtype=np.int32)
    keep, pointers, sizes, poses, ranges = _cloud_arrays(clouds)
    code = _voxelizer_lib().vgt_ref_voxelize_f64(
        occ.ctypes.data_as(_f32p), *occ.shape, float(voxel_size), len(clouds), pointers, sizes,
        poses.ctypes.data_as(_f64p), ranges.ctypes.data_as(_f64p), float(percent_seen_free),
        int(outlier_points_threshold), int(num_cameras_seen_free), threads,
        out.ctypes.data_as(_f32p), counts.ctypes.data_as(_i32p))
    del keep
    if code != 0:
        raise RuntimeError("reference voxelizer failed")
    return out, counts


def voxelize_posed(static_occupancy, x_wg, clouds, voxel_size: float, percent_seen_free: float,
                   outlier_points_threshold: int, num_cameras_seen_free: int, threads: int = 0):
    """The interface call with a posed grid: clouds = [(points, x_wc, max_range), ...]; the
    reference composes X_GC = X_WG^-1 * X_WC itself. Returns the filtered occupancy."""
    occ = np.ascontiguousarray(static_occupancy, dtype=np.float32)
    out = np.empty_like(occ)
    keep, pointers, sizes, poses, ranges = _cloud_arrays(clouds)
    origin = np.ascontiguousarray(np.asarray(x_wg, dtype=np.float64).reshape(4, 4).T)
    code = _voxelizer_lib().vgt_ref_voxelize_posed_f64(
        occ.ctypes.data_as(_f32p), *occ.shape, float(voxel_size), origin.ctypes.data_as(_f64p),
        len(clouds), pointers, sizes, poses.ctypes.data_as(_f64p), ranges.ctypes.data_as(_f64p),
        float(percent_seen_free), int(outlier_points_threshold), int(num_cameras_seen_free),
        threads, out.ctypes.data_as(_f32p))
    del keep
    if code != 0:
        raise RuntimeError("reference voxelizer failed")
    return out


def raycast_single(origin, point, max_range: float, dims, voxel_size: float):
    nx, ny, nz = (int(d) for d in dims)
    counts = np.zeros((nx, ny, nz, 2), dtype=np.int32)
    o = np.ascontiguousarray(origin, dtype=np.float64)
    p = np.ascontiguousarray(point, dtype=np.float64)
    if _voxelizer_lib().vgt_ref_raycast_single_f64(
            o.ctypes.data_as(_f64p), p.ctypes.data_as(_f64p), float(max_range), nx, ny, nz,
            float(voxel_size), counts.ctypes.data_as(_i32p)):
        raise ValueError("reference RaycastSinglePoint threw")
    return counts


def isometry_product(a, b) -> np.ndarray:
    """a * b for rigid 4x4 transforms (row-major numpy in and out), in the fixed operation order
    of the stand-in (oracle/ref_shim/Eigen/Geometry)."""
    a_cm = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(4, 4).T)
    b_cm = np.ascontiguousarray(np.asarray(b, dtype=np.float64).reshape(4, 4).T)
    out = np.empty((4, 4), dtype=np.float64)
    _voxelizer_lib().vgt_ref_isometry_product(
        a_cm.ctypes.data_as(_f64p), b_cm.ctypes.data_as(_f64p), out.ctypes.data_as(_f64p))
    return np.ascontiguousarray(out.T)


def isometry_inverse(a) -> np.ndarray:
    a_cm = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(4, 4).T)
    out = np.empty((4, 4), dtype=np.float64)
    _voxelizer_lib().vgt_ref_isometry_inverse(a_cm.ctypes.data_as(_f64p),
                                              out.ctypes.data_as(_f64p))
    return np.ascontiguousarray(out.T)


# ---- the reference's own SignedDistanceField query members (ref_shim/ref_sdf_queries_entry.cpp)
QUERY_ESTIMATE_DISTANCE, QUERY_COARSE_GRADIENT, QUERY_FINE_GRADIENT, QUERY_PROJECT = 0, 1, 2, 3


def _origin_column_major(origin_transform):
    origin = np.eye(4) if origin_transform is None else np.asarray(origin_transform, np.float64)
    return np.ascontiguousarray(origin.T).reshape(16)      # column-major = rows of the transpose


def sdf_query(field, resolution: float, origin_transform, kind: int, points, a=0.0, b=0.0):
    """(values [count, 1 or 3], status [count]: 0 no value, 1 value, 2 the reference threw)."""
    handle = lib()
    handle.vgt_ref_sdf_query.argtypes = [
        _f32p, _i64, _i64, _i64, ctypes.c_double, _f64p, _int, ctypes.c_double, ctypes.c_double,
        _f64p, _i64, _f64p, ctypes.POINTER(ctypes.c_uint8)]
    field = np.ascontiguousarray(field, dtype=np.float32)
    points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    origin = _origin_column_major(origin_transform)
    width = 1 if kind == QUERY_ESTIMATE_DISTANCE else 3
    values = np.zeros((len(points), width), dtype=np.float64)
    status = np.zeros(len(points), dtype=np.uint8)
    code = handle.vgt_ref_sdf_query(
        field.ctypes.data_as(_f32p), *field.shape, float(resolution),
        origin.ctypes.data_as(_f64p), int(kind), float(a), float(b),
        points.ctypes.data_as(_f64p), len(points), values.ctypes.data_as(_f64p),
        status.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    if code != 0:
        raise RuntimeError("reference SDF query failed")
    return values, status


def sdf_local_extrema_map(field, resolution: float, origin_transform=None) -> np.ndarray:
    handle = lib()
    handle.vgt_ref_sdf_local_extrema_map.argtypes = [
        _f32p, _i64, _i64, _i64, ctypes.c_double, _f64p, _f64p]
    field = np.ascontiguousarray(field, dtype=np.float32)
    origin = _origin_column_major(origin_transform)
    out = np.empty(field.shape + (3,), dtype=np.float64)
    code = handle.vgt_ref_sdf_local_extrema_map(
        field.ctypes.data_as(_f32p), *field.shape, float(resolution),
        origin.ctypes.data_as(_f64p), out.ctypes.data_as(_f64p))
    if code != 0:
        raise RuntimeError("reference ComputeLocalExtremaMap failed")
    return out


# ---- the reference's own mesh rasterizer (ref_shim/ref_mesh_rasterizer_entry.cpp)
def rasterize_mesh(vertices, triangles, occupancy, resolution: float, origin_transform=None,
                   enforce_contains: bool = False, threads: int = 1) -> int:
    handle = lib()
    handle.vgt_ref_rasterize_mesh.argtypes = [
        _f64p, _i64, ctypes.POINTER(ctypes.c_int32), _i64, _f32p, _i64, _i64, _i64,
        ctypes.c_double, _f64p, _int, _int]
    vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
    triangles = np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 3)
    assert occupancy.dtype == np.float32 and occupancy.flags.c_contiguous
    origin = _origin_column_major(origin_transform)
    return int(handle.vgt_ref_rasterize_mesh(
        vertices.ctypes.data_as(_f64p), len(vertices),
        triangles.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(triangles),
        occupancy.ctypes.data_as(_f32p), *occupancy.shape, float(resolution),
        origin.ctypes.data_as(_f64p), int(enforce_contains), int(threads)))


def rasterize_mesh_into_occupancy_map(vertices, triangles, resolution: float, threads: int = 1):
    """(occupancy [nx, ny, nz], origin 4x4, status) of RasterizeMeshIntoOccupancyMap."""
    handle = lib()
    handle.vgt_ref_rasterize_mesh_into_map.argtypes = [
        _f64p, _i64, ctypes.POINTER(ctypes.c_int32), _i64, ctypes.c_double, _int,
        ctypes.POINTER(_i64), _f64p, _f32p, _i64]
    vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
    triangles = np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 3)
    dims = (_i64 * 3)()
    origin = np.zeros(16, dtype=np.float64)
    args = (vertices.ctypes.data_as(_f64p), len(vertices),
            triangles.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(triangles),
            float(resolution), int(threads), dims, origin.ctypes.data_as(_f64p))
    code = int(handle.vgt_ref_rasterize_mesh_into_map(*args, None, 0))
    if code != 0:
        return None, None, code
    occupancy = np.zeros(tuple(int(d) for d in dims), dtype=np.float32)
    code = int(handle.vgt_ref_rasterize_mesh_into_map(
        *args, occupancy.ctypes.data_as(_f32p), occupancy.size))
    return occupancy, origin.reshape(4, 4).T.copy(), code
