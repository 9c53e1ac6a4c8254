"""Host-side mirror of the point cloud voxelization interface for the B200 backend.

Reference interfaces mirrored (paths relative to the reference checkout):

* ``SeenAs`` / ``PointCloudVoxelizationFilterOptions``  include/.../pointcloud_voxelization_interface.hpp:18-92
* ``PointCloudWrapper``                                 pointcloud_voxelization_interface.hpp:94-202
* ``VoxelizerRuntime``                                  pointcloud_voxelization_interface.hpp:206-229
* ``PointCloudVoxelizationInterface.VoxelizePointClouds``  pointcloud_voxelization_interface.hpp:246-292
* ``B200PointCloudVoxelizer``  takes the place of ``CudaPointCloudVoxelizer``
  (include/.../device_pointcloud_voxelization.hpp:67-73); unlike it, this backend raycasts in
  double precision so its counts equal the CPU backend's (src/.../cpu_pointcloud_voxelization.cpp).
"""
from __future__ import annotations

import ctypes
import enum
import math
from typing import Callable, Sequence

import numpy as np

from . import _capi
from .grids import OccupancyMap, compose_rigid


class SeenAs(enum.IntEnum):
    UNKNOWN = 0
    FILLED = 1
    FREE = 2


class PointCloudVoxelizationFilterOptions:
    def __init__(self, percent_seen_free: float = 1.0, outlier_points_threshold: int = 1,
                 num_cameras_seen_free: int = 1):
        # Same checks and messages as pointcloud_voxelization_interface.hpp:30-41.
        if percent_seen_free <= 0.0 or percent_seen_free > 1.0 or math.isnan(percent_seen_free):
            raise ValueError("0 < percent_seen_free_ <= 1 must be true")
        if outlier_points_threshold <= 0:
            raise ValueError("outlier_points_threshold_ <= 0")
        if num_cameras_seen_free <= 0:
            raise ValueError("num_cameras_seen_free_ <= 0")
        self._percent_seen_free = float(percent_seen_free)
        self._outlier_points_threshold = int(outlier_points_threshold)
        self._num_cameras_seen_free = int(num_cameras_seen_free)

    def PercentSeenFree(self) -> float:
        return self._percent_seen_free

    def OutlierPointsThreshold(self) -> int:
        return self._outlier_points_threshold

    def NumCamerasSeenFree(self) -> int:
        return self._num_cameras_seen_free

    def CountsSeenAs(self, seen_free_count: int, seen_filled_count: int) -> SeenAs:
        """pointcloud_voxelization_interface.hpp:55-86 (host restatement for callers; the device
        filter kernel applies the same rule)."""
        filtered = seen_filled_count if seen_filled_count >= self._outlier_points_threshold else 0
        if seen_free_count > 0 and filtered > 0:
            fraction = float(seen_free_count) / float(seen_free_count + filtered)
            return SeenAs.FREE if fraction >= self._percent_seen_free else SeenAs.FILLED
        if seen_free_count > 0:
            return SeenAs.FREE
        if filtered > 0:
            return SeenAs.FILLED
        return SeenAs.UNKNOWN

    def as_struct(self) -> _capi.FilterOptions:
        return _capi.FilterOptions(self._percent_seen_free, self._outlier_points_threshold,
                                   self._num_cameras_seen_free)


class PointCloudWrapper:
    """Abstract cloud (pointcloud_voxelization_interface.hpp:94-202)."""

    def MaxRange(self) -> float:
        raise NotImplementedError

    def Size(self) -> int:
        raise NotImplementedError

    def PointCloudOriginTransform(self) -> np.ndarray:
        raise NotImplementedError

    def CopyPointLocationIntoDoublePtrImpl(self, point_index: int, destination: np.ndarray):
        raise NotImplementedError

    def EnforcePointIndexInRange(self, point_index: int) -> None:
        if point_index < 0 or point_index >= self.Size():
            raise IndexError("point_index out of range")

    def GetPointLocationVector4d(self, point_index: int) -> np.ndarray:
        self.EnforcePointIndexInRange(point_index)
        point = np.array([0.0, 0.0, 0.0, 1.0])
        self.CopyPointLocationIntoDoublePtrImpl(point_index, point[:3])
        return point

    def PointsAsDoubleArray(self) -> np.ndarray:
        """All points as a contiguous float64 [N, 3] array. Subclasses holding a dense buffer
        override this to avoid the per-point virtual call the reference pays
        (device_pointcloud_voxelization.cpp:130-136)."""
        points = np.empty((self.Size(), 3), dtype=np.float64)
        for index in range(self.Size()):
            self.CopyPointLocationIntoDoublePtrImpl(index, points[index])
        return points


class VectorPointCloudWrapper(PointCloudWrapper):
    """Dense [N, 3] float64 cloud; the analogue of the test fixture
    VectorVector3dPointCloudWrapper (test/pointcloud_voxelization_test.cpp:31-82)."""

    def __init__(self, points=None, origin_transform=None, max_range: float = float("inf")):
        self._points = (np.zeros((0, 3), dtype=np.float64) if points is None
                        else np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3))
        self._origin_transform = (np.eye(4) if origin_transform is None
                                  else np.array(origin_transform, dtype=np.float64).reshape(4, 4))
        self._max_range = float(max_range)

    def PushBack(self, point) -> None:
        self._points = np.vstack([self._points, np.asarray(point, dtype=np.float64).reshape(1, 3)])

    def MaxRange(self) -> float:
        return self._max_range

    def SetMaxRange(self, max_range: float) -> None:
        self._max_range = float(max_range)

    def Size(self) -> int:
        return int(self._points.shape[0])

    def PointCloudOriginTransform(self) -> np.ndarray:
        return self._origin_transform

    def SetPointCloudOriginTransform(self, origin_transform) -> None:
        self._origin_transform = np.array(origin_transform, dtype=np.float64).reshape(4, 4)

    def CopyPointLocationIntoDoublePtrImpl(self, point_index: int, destination: np.ndarray):
        destination[:3] = self._points[point_index]

    def PointsAsDoubleArray(self) -> np.ndarray:
        return self._points


class Float32PointCloudWrapper(VectorPointCloudWrapper):
    """Dense [N, 3] float32 cloud: what a sensor_msgs/PointCloud2 carries. The reference's
    PointCloud2Wrapper (pointcloud_voxelization_ros_interface.hpp:35-97) widens each point to
    double when it is asked for one; here the float32 buffer goes to the device as it is (half
    the bytes) and is widened there (vgt_b200_voxelize_f32). Same counts as the same values held
    in a VectorPointCloudWrapper."""

    def __init__(self, points=None, origin_transform=None, max_range: float = float("inf")):
        super().__init__(None, origin_transform, max_range)
        self._points32 = (np.zeros((0, 3), dtype=np.float32) if points is None
                          else np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3))

    def Size(self) -> int:
        return int(self._points32.shape[0])

    def CopyPointLocationIntoDoublePtrImpl(self, point_index: int, destination: np.ndarray):
        destination[:3] = self._points32[point_index]

    def PointsAsDoubleArray(self) -> np.ndarray:
        return self._points32.astype(np.float64)

    def PointsAsFloatArray(self) -> np.ndarray:
        return self._points32


class VoxelizerRuntime:
    def __init__(self, raycasting_time: float, filtering_time: float):
        if raycasting_time < 0.0:
            raise ValueError("raycasting_time < 0.0")
        if filtering_time < 0.0:
            raise ValueError("filtering_time < 0.0")
        self._raycasting_time = raycasting_time
        self._filtering_time = filtering_time

    def RaycastingTime(self) -> float:
        return self._raycasting_time

    def FilteringTime(self) -> float:
        return self._filtering_time


def grid_from_cloud_transform(static_environment: OccupancyMap,
                              cloud: PointCloudWrapper) -> np.ndarray:
    """X_GC = X_GW * X_WC (cpu_pointcloud_voxelization.cpp:172-176)."""
    return compose_rigid(static_environment.InverseOriginTransform(),
                         cloud.PointCloudOriginTransform())


class PointCloudVoxelizationInterface:
    """pointcloud_voxelization_interface.hpp:231-302: argument checks, then the backend."""

    def VoxelizePointClouds(self, static_environment: OccupancyMap,
                            filter_options: PointCloudVoxelizationFilterOptions,
                            pointclouds: Sequence[PointCloudWrapper],
                            runtime_log_fn: Callable[[VoxelizerRuntime], None] | None = None,
                            output_environment: OccupancyMap | None = None) -> OccupancyMap:
        if static_environment is None or not static_environment.IsInitialized():
            raise ValueError("!static_environment.IsInitialized()")
        if output_environment is None:
            output_environment = static_environment.copy()
        elif output_environment.ControlSizes() != static_environment.ControlSizes():
            raise ValueError(
                "static_environment.ControlSizes() != output_environment.ControlSizes()")
        for index, cloud in enumerate(pointclouds):
            if cloud is None:
                raise ValueError(f"pointclouds[{index}] is null")
        runtime = self.DoVoxelizePointClouds(static_environment, filter_options, pointclouds,
                                             output_environment)
        if runtime_log_fn:
            runtime_log_fn(runtime)
        return output_environment

    def DoVoxelizePointClouds(self, static_environment, filter_options, pointclouds,
                              output_environment) -> VoxelizerRuntime:
        raise NotImplementedError


class B200PointCloudVoxelizer(PointCloudVoxelizationInterface):
    """The sm_100a backend. Options follow the reference's string -> int32 map
    (src/.../cuda_voxelization_helpers.cu:566-590): ``CUDA_DEVICE`` selects the GPU; the CPU thread
    options of the other backends are accepted and ignored."""

    def __init__(self, options: dict | None = None, logging_fn: Callable[[str], None] | None = None):
        options = dict(options or {})
        self._device = int(options.get("CUDA_DEVICE", 0))
        if logging_fn:
            logging_fn(f"Option [CUDA_DEVICE] = {self._device}")
        # EnforceAvailable (device_pointcloud_voxelization.hpp:34-46): unusable -> runtime_error.
        _capi.require_device(self._device)
        self.last_counts = None

    def DoVoxelizePointClouds(self, static_environment, filter_options, pointclouds,
                              output_environment, keep_counts: bool = False) -> VoxelizerRuntime:
        lib = _capi.library()
        nx, ny, nz = static_environment.ControlSizes().shape
        # float32 clouds only: the float32 entry (12 bytes per point over the bus)
        single = len(pointclouds) > 0 and all(hasattr(cloud, "PointsAsFloatArray")
                                              for cloud in pointclouds)
        cloud_type, scalar, c_scalar, entry = (
            (_capi.CloudF32, np.float32, ctypes.c_float, lib.vgt_b200_voxelize_f32) if single
            else (_capi.Cloud, np.float64, ctypes.c_double, lib.vgt_b200_voxelize_f64))
        clouds = (cloud_type * max(1, len(pointclouds)))()
        keepalive = []
        for index, cloud in enumerate(pointclouds):
            points = np.ascontiguousarray(
                cloud.PointsAsFloatArray() if single else cloud.PointsAsDoubleArray(),
                dtype=scalar)
            keepalive.append(points)
            x_gc = grid_from_cloud_transform(static_environment, cloud)
            clouds[index].points_xyz = points.ctypes.data_as(ctypes.POINTER(c_scalar))
            clouds[index].num_points = points.shape[0]
            clouds[index].x_gc = (ctypes.c_double * 16)(*x_gc.T.reshape(-1))  # column-major
            clouds[index].max_range = float(cloud.MaxRange())
        options = filter_options.as_struct()
        counts = None
        if keep_counts and len(pointclouds) > 0:
            counts = np.empty((len(pointclouds), nx, ny, nz, 2), dtype=np.int32)
        seconds = (ctypes.c_double * 2)()
        static_data = static_environment.GetImmutableRawData()
        out_data = output_environment.GetMutableRawData()
        code = entry(
            static_data.ctypes.data, nx, ny, nz, static_environment.VoxelXSize(), clouds,
            len(pointclouds), ctypes.byref(options), self._device, out_data.ctypes.data,
            None if counts is None else counts.ctypes.data, seconds)
        _capi.check(code)
        self.last_counts = counts
        return VoxelizerRuntime(seconds[0], seconds[1])

    def VoxelizePointCloudsWithCounts(self, static_environment, filter_options, pointclouds):
        """Test / diagnostics helper: filtered map plus per-cloud int32 [x, y, z, 2] counts."""
        output = static_environment.copy()
        for index, cloud in enumerate(pointclouds):
            if cloud is None:
                raise ValueError(f"pointclouds[{index}] is null")
        self.DoVoxelizePointClouds(static_environment, filter_options, pointclouds, output,
                                   keep_counts=True)
        counts = self.last_counts
        if counts is None:
            counts = np.zeros((0,) + static_environment.ControlSizes().shape + (2,), np.int32)
        return output, counts


def GetAvailableBackends():
    """pointcloud_voxelization.hpp:54-68 analogue: one entry per usable B200, nothing else
    (no multi-backend dispatch, no CPU fallback)."""
    return [{"device_name": f"B200 device {index}", "options": {"CUDA_DEVICE": index}}
            for index in range(_capi.device_count())]


def MakePointCloudVoxelizer(options: dict | None = None, logging_fn=None):
    return B200PointCloudVoxelizer(options, logging_fn)
