"""ctypes front-end for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and
``__graft_entry__.smoke()`` may import this module. The product package
``voxelized_geometry_tools_b200`` never does (tests/test_no_oracle_in_product.py
enforces it).

Two libraries can sit behind it:

* ``oracle/_build/libvgt_oracle.so`` -- our restatement (``edt_oracle.cpp``,
  ``voxelizer_oracle.cpp``), rebuilt on the current host when missing or when the
  host CPU differs from the one it was built on (it is compiled ``-march=native``
  like the reference, CMakeLists.txt.ros2:58,70).
* ``oracle/_ref/libvgt_ref.so`` -- the reference's own
  ``signed_distance_field_generation.cpp`` compiled unmodified over shim headers
  (``make -C oracle ref``; needs ``/root/reference``, so it is built in the dev
  container and travels to the GPU box prebuilt).
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libvgt_oracle.so"
_STAMP_PATH = _HERE / "_build" / "host.stamp"
_REF_PATH = _HERE / "_ref" / "libvgt_ref.so"

_c_f32p = ctypes.POINTER(ctypes.c_float)
_c_f64p = ctypes.POINTER(ctypes.c_double)
_c_i32p = ctypes.POINTER(ctypes.c_int32)
_c_u8p = ctypes.POINTER(ctypes.c_uint8)
_i64 = ctypes.c_int64
_int = ctypes.c_int


def _host_signature() -> str:
    try:
        with open("/proc/cpuinfo", "r", encoding="utf-8") as handle:
            for line in handle:
                if line.startswith("flags"):
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def _sources_signature() -> str:
    digest = hashlib.sha1()
    for name in ("edt_oracle.cpp", "voxelizer_oracle.cpp", "mesh_rasterizer_oracle.cpp",
                 "Makefile"):
        digest.update((_HERE / name).read_bytes())
    return digest.hexdigest()


def build(force: bool = False) -> Path:
    """Compiles the restated oracle for this host if needed; returns the .so path."""
    stamp = _host_signature() + ":" + _sources_signature()
    if (not force and _LIB_PATH.exists() and _STAMP_PATH.exists()
            and _STAMP_PATH.read_text().strip() == stamp):
        return _LIB_PATH
    env = dict(os.environ)
    subprocess.run(["make", "-C", str(_HERE), "-B", "_build/libvgt_oracle.so"],
                   check=True, env=env, capture_output=True)
    _STAMP_PATH.write_text(stamp)
    return _LIB_PATH


def build_reference(force: bool = False) -> Path | None:
    """Compiles the reference's own EDT source (needs /root/reference). None if absent."""
    maps_path = _REF_PATH.with_name("libvgt_ref_maps.so")
    if _REF_PATH.exists() and maps_path.exists() and not force:
        return _REF_PATH
    if not Path("/root/reference/src/voxelized_geometry_tools"
                "/signed_distance_field_generation.cpp").exists():
        return _REF_PATH if _REF_PATH.exists() else None
    subprocess.run(["make", "-C", str(_HERE), "-B", "ref"], check=True,
                   capture_output=True)
    return _REF_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        handle = ctypes.CDLL(str(build()))
        handle.vgt_oracle_max_threads.restype = _int
        handle.vgt_oracle_edt_sq_f64.argtypes = [
            _c_f32p, _i64, _i64, _i64, _int, _int, _c_f64p, _c_f64p]
        handle.vgt_oracle_transform_inplace_f64.argtypes = [
            _c_f64p, _i64, _i64, _i64, _int]
        handle.vgt_oracle_sdf_f32.argtypes = [
            _c_f32p, _i64, _i64, _i64, ctypes.c_double, _int, _int, _int,
            _c_f32p, _c_f32p]
        handle.vgt_oracle_sdf_f64.argtypes = [
            _c_f32p, _i64, _i64, _i64, ctypes.c_double, _int, _int, _int,
            _c_f64p, _c_f64p]
        handle.vgt_oracle_sdf_from_mask_f32.argtypes = [
            _c_u8p, _i64, _i64, _i64, ctypes.c_double, _int, _int, _c_f32p,
            _c_f32p]
        handle.vgt_oracle_raycast_cloud_f64.argtypes = [
            _c_f64p, _i64, _c_f64p, ctypes.c_double, _i64, _i64, _i64,
            ctypes.c_double, _int, _c_i32p]
        handle.vgt_oracle_raycast_single_f64.argtypes = [
            _c_f64p, _c_f64p, ctypes.c_double, _i64, _i64, _i64,
            ctypes.c_double, _c_i32p]
        handle.vgt_oracle_filter_f32.argtypes = [
            _c_i32p, ctypes.c_int32, _i64, ctypes.c_double, ctypes.c_int32,
            ctypes.c_int32, _int, _c_f32p]
        handle.vgt_oracle_rasterize_mesh_f64.argtypes = [
            _c_f64p, _i64, _c_i32p, _i64, _c_f32p, _i64, _i64, _i64, ctypes.c_double, _c_f64p,
            _c_f64p, _int]
        handle.vgt_oracle_mesh_map_extent_f64.argtypes = [
            _c_f64p, _i64, ctypes.c_double, ctypes.POINTER(_i64), _c_f64p]
        _lib = handle
    return _lib


def max_threads() -> int:
    return int(lib().vgt_oracle_max_threads())


def _occ3(occupancy) -> np.ndarray:
    occ = np.ascontiguousarray(occupancy, dtype=np.float32)
    if occ.ndim != 3:
        raise ValueError("occupancy must be a 3-D array indexed [x, y, z]")
    return occ


def _check(code: int, what: str) -> None:
    if code != 0:
        raise RuntimeError(f"oracle {what} failed with code {code}")


def edt_squared(occupancy, unknown_is_filled: bool = True, threads: int = 0):
    """(dist_to_filled_sq, dist_to_free_sq) as float64 arrays in voxel units, inf if none."""
    occ = _occ3(occupancy)
    to_filled = np.empty(occ.shape, dtype=np.float64)
    to_free = np.empty(occ.shape, dtype=np.float64)
    _check(lib().vgt_oracle_edt_sq_f64(
        occ.ctypes.data_as(_c_f32p), *occ.shape, int(unknown_is_filled),
        threads, to_filled.ctypes.data_as(_c_f64p),
        to_free.ctypes.data_as(_c_f64p)), "edt_sq")
    return to_filled, to_free


def transform_inplace(field: np.ndarray, threads: int = 0) -> np.ndarray:
    """ComputeDistanceFieldTransformInPlace on a float64 [x, y, z] field."""
    assert field.dtype == np.float64 and field.flags.c_contiguous
    _check(lib().vgt_oracle_transform_inplace_f64(
        field.ctypes.data_as(_c_f64p), *field.shape, threads), "transform")
    return field


def sdf(occupancy, resolution: float, unknown_is_filled: bool = True,
        add_virtual_border: bool = False, threads: int = 0,
        dtype=np.float32):
    """ExtractSignedDistanceField<dtype>; returns (sdf[x,y,z], (min, max))."""
    occ = _occ3(occupancy)
    out = np.empty(occ.shape, dtype=dtype)
    min_max = np.zeros(2, dtype=dtype)
    if np.dtype(dtype) == np.float32:
        code = lib().vgt_oracle_sdf_f32(
            occ.ctypes.data_as(_c_f32p), *occ.shape, float(resolution),
            int(unknown_is_filled), int(add_virtual_border), threads,
            out.ctypes.data_as(_c_f32p), min_max.ctypes.data_as(_c_f32p))
    elif np.dtype(dtype) == np.float64:
        code = lib().vgt_oracle_sdf_f64(
            occ.ctypes.data_as(_c_f32p), *occ.shape, float(resolution),
            int(unknown_is_filled), int(add_virtual_border), threads,
            out.ctypes.data_as(_c_f64p), min_max.ctypes.data_as(_c_f64p))
    else:
        raise ValueError("dtype must be float32 or float64")
    _check(code, "sdf")
    return out, (min_max[0], min_max[1])


def sdf_from_mask(mask, resolution: float, add_virtual_border: bool = False,
                  threads: int = 0):
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty(m.shape, dtype=np.float32)
    min_max = np.zeros(2, dtype=np.float32)
    _check(lib().vgt_oracle_sdf_from_mask_f32(
        m.ctypes.data_as(_c_u8p), *m.shape, float(resolution),
        int(add_virtual_border), threads, out.ctypes.data_as(_c_f32p),
        min_max.ctypes.data_as(_c_f32p)), "sdf_from_mask")
    return out, (min_max[0], min_max[1])


# ------------------------------------------------------------------------------------------------
# The other map types. Plain numpy over the pinned pieces above:
# the only new arithmetic is the filled predicate and the free / named merge.
# ------------------------------------------------------------------------------------------------
def cells_filled_mask(cells, unknown_is_filled: bool = True, objects_to_use=(),
                      named_only: bool = False):
    """The is_filled_fn of the cell maps, evaluated over the whole grid:
    occupancy_component_map.hpp:276-299 (occupancy rule only),
    tagged_object_occupancy_map.hpp:213-241 (occupancy rule and: no objects listed, or the
    cell's object id is listed), tagged_object_occupancy_map.hpp:316-334 (named objects: id > 0)."""
    occupancy = cells["occupancy"]
    filled = occupancy > np.float32(0.5)
    if unknown_is_filled:
        filled = filled | (occupancy == np.float32(0.5))
    if named_only:
        filled = filled & (cells["object_id"] > 0)
    else:
        listed = [int(i) for i in objects_to_use]
        if listed:
            filled = filled & np.isin(cells["object_id"], np.array(listed, dtype=np.uint32))
    return filled


def sdf_from_cells(cells, resolution: float, unknown_is_filled: bool = True,
                   add_virtual_border: bool = False, objects_to_use=(), named_only: bool = False,
                   dtype=np.float32, threads: int = 0):
    """ExtractSignedDistanceField of a cell map: predicate -> internal::ExtractSignedDistanceField
    (signed_distance_field_generation.hpp:115-285). Returns (sdf, (min, max))."""
    mask = cells_filled_mask(cells, unknown_is_filled, objects_to_use, named_only)
    # an occupancy of exactly 1 / 0 makes the OccupancyMap entry evaluate the same predicate
    return sdf(mask.astype(np.float32), resolution, True, add_virtual_border, threads, dtype)


def sdf_free_and_named(cells, resolution: float, unknown_is_filled: bool = True,
                       add_virtual_border: bool = False, dtype=np.float32, threads: int = 0):
    """ExtractFreeAndNamedObjectsSignedDistanceField (tagged_object_occupancy_map.hpp:293-378):
    free >= 0 -> free; else named <= -0 -> named; else 0; then Lock()'s min/max."""
    free, _ = sdf_from_cells(cells, resolution, unknown_is_filled, add_virtual_border, (),
                             False, dtype, threads)
    named, _ = sdf_from_cells(cells, resolution, unknown_is_filled, add_virtual_border, (),
                              True, dtype, threads)
    zero = np.zeros((), dtype=dtype)
    combined = np.where(free >= zero, free, np.where(named <= -zero, named, zero))
    combined = combined.astype(dtype)
    return combined, (combined.min(), combined.max())


def raycast_cloud(points_xyz, x_gc, max_range: float, dims, voxel_size: float,
                  counts: np.ndarray | None = None, threads: int = 0):
    """Accumulates one cloud into counts[x, y, z, 2] (0 = seen free, 1 = seen filled).

    ``x_gc`` is a 4x4 grid-from-cloud matrix (row-major numpy, as written on paper).
    """
    pts = np.ascontiguousarray(points_xyz, dtype=np.float64).reshape(-1, 3)
    xgc = np.asarray(x_gc, dtype=np.float64).reshape(4, 4)
    column_major = np.ascontiguousarray(xgc.T).reshape(-1)
    nx, ny, nz = (int(d) for d in dims)
    if counts is None:
        counts = np.zeros((nx, ny, nz, 2), dtype=np.int32)
    assert counts.dtype == np.int32 and counts.flags.c_contiguous
    _check(lib().vgt_oracle_raycast_cloud_f64(
        pts.ctypes.data_as(_c_f64p), pts.shape[0],
        column_major.ctypes.data_as(_c_f64p), float(max_range), nx, ny, nz,
        float(voxel_size), threads, counts.ctypes.data_as(_c_i32p)), "raycast")
    return counts


def raycast_single(origin, point, max_range: float, dims, voxel_size: float,
                   counts: np.ndarray | None = None):
    nx, ny, nz = (int(d) for d in dims)
    if counts is None:
        counts = np.zeros((nx, ny, nz, 2), dtype=np.int32)
    o = np.ascontiguousarray(origin, dtype=np.float64)
    p = np.ascontiguousarray(point, dtype=np.float64)
    code = lib().vgt_oracle_raycast_single_f64(
        o.ctypes.data_as(_c_f64p), p.ctypes.data_as(_c_f64p), float(max_range),
        nx, ny, nz, float(voxel_size), counts.ctypes.data_as(_c_i32p))
    if code == 2:
        raise ValueError("non-finite origin or point")
    _check(code, "raycast_single")
    return counts


def filter_grids(counts, occupancy, percent_seen_free: float,
                 outlier_points_threshold: int, num_cameras_seen_free: int,
                 threads: int = 0):
    """counts[grid, x, y, z, 2] + static occupancy[x, y, z] -> filtered occupancy (copy)."""
    c = np.ascontiguousarray(counts, dtype=np.int32)
    occ = np.array(occupancy, dtype=np.float32, copy=True, order="C")
    num_grids = 0 if c.size == 0 else c.shape[0]
    code = lib().vgt_oracle_filter_f32(
        c.ctypes.data_as(_c_i32p), num_grids, occ.size,
        float(percent_seen_free), int(outlier_points_threshold),
        int(num_cameras_seen_free), threads, occ.ctypes.data_as(_c_f32p))
    if code == 2:
        raise ValueError("invalid filter options")
    _check(code, "filter")
    return occ


def voxelize(static_occupancy, clouds, voxel_size: float, percent_seen_free: float,
             outlier_points_threshold: int, num_cameras_seen_free: int,
             threads: int = 0):
    """Whole CPU voxelizer: clouds = [(points_xyz, x_gc 4x4, max_range), ...].

    Returns (filtered occupancy, counts[grid, x, y, z, 2]). Mirrors
    CpuPointCloudVoxelizer::DoVoxelizePointClouds (cpu_pcv.cpp:133-165).
    """
    occ = _occ3(static_occupancy)
    dims = occ.shape
    counts = np.zeros((len(clouds),) + tuple(dims) + (2,), dtype=np.int32)
    for index, (points, x_gc, max_range) in enumerate(clouds):
        raycast_cloud(points, x_gc, max_range, dims, voxel_size,
                      counts=counts[index], threads=threads)
    filtered = filter_grids(counts, occ, percent_seen_free,
                            outlier_points_threshold, num_cameras_seen_free,
                            threads)
    return filtered, counts


# ---- mesh rasterizer (mesh_rasterizer_oracle.cpp)
RASTERIZE_OK, RASTERIZE_NOT_CONTAINED, RASTERIZE_BAD_INDEX = 0, 2, 3


def _inverse_rigid_fixed_order(m):
    """Inverse of a rigid 4x4 in the stand-in Eigen's order (ref_shim/Eigen/Geometry)."""
    m = np.asarray(m, dtype=np.float64)
    out = np.eye(4)
    out[:3, :3] = m[:3, :3].T
    for r in range(3):
        out[r, 3] = -((out[r, 0] * m[0, 3] + out[r, 1] * m[1, 3]) + out[r, 2] * m[2, 3])
    return out


def rasterize_mesh(vertices, triangles, occupancy, resolution: float, origin_transform=None,
                   enforce_contains: bool = False):
    """Sets the cells of `occupancy` (float32 [nx, ny, nz], modified in place) the triangles
    touch to 1.0; returns the status code (RASTERIZE_*)."""
    vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
    triangles = np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 3)
    assert occupancy.dtype == np.float32 and occupancy.flags.c_contiguous
    origin = np.eye(4) if origin_transform is None else np.asarray(origin_transform, np.float64)
    x_wg = np.ascontiguousarray(origin.T).reshape(-1)
    x_gw = np.ascontiguousarray(_inverse_rigid_fixed_order(origin).T).reshape(-1)
    return int(lib().vgt_oracle_rasterize_mesh_f64(
        vertices.ctypes.data_as(_c_f64p), len(vertices), triangles.ctypes.data_as(_c_i32p),
        len(triangles), occupancy.ctypes.data_as(_c_f32p), *occupancy.shape, float(resolution),
        x_wg.ctypes.data_as(_c_f64p), x_gw.ctypes.data_as(_c_f64p), int(enforce_contains)))


def mesh_map_extent(vertices, resolution: float):
    """(dims, origin translation) of the map RasterizeMeshIntoOccupancyMap builds."""
    vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
    dims = (_i64 * 3)()
    translation = np.zeros(3, dtype=np.float64)
    _check(lib().vgt_oracle_mesh_map_extent_f64(
        vertices.ctypes.data_as(_c_f64p), len(vertices), float(resolution), dims,
        translation.ctypes.data_as(_c_f64p)), "mesh map extent")
    return tuple(int(d) for d in dims), translation


def rasterize_mesh_into_occupancy_map(vertices, triangles, resolution: float):
    dims, translation = mesh_map_extent(vertices, resolution)
    origin = np.eye(4)
    origin[:3, 3] = translation
    occupancy = np.zeros(dims, dtype=np.float32)
    code = rasterize_mesh(vertices, triangles, occupancy, resolution, origin, True)
    return occupancy, origin, code
