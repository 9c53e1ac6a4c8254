"""Device-resident entry points: torch CUDA tensors in, torch CUDA tensors out.

torch is only the owner of device memory and streams here; every kernel is launched by
libvgt_b200.so through the ``*_dev`` C-ABI functions on torch's current stream.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _capi
from .pointcloud_voxelization import PointCloudVoxelizationFilterOptions


def _stream_handle(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(tensor: torch.Tensor, dtype: torch.dtype, what: str) -> None:
    if not tensor.is_cuda:
        raise _capi.BackendUnavailable(f"{what} must live on a CUDA device (no CPU fallback)")
    if tensor.dtype != dtype or not tensor.is_contiguous():
        raise ValueError(f"{what} must be a contiguous {dtype} tensor")


def signed_distance_field(occupancy: torch.Tensor, resolution: float,
                          unknown_is_filled: bool = True, add_virtual_border: bool = False,
                          out: torch.Tensor | None = None, min_max: torch.Tensor | None = None,
                          compute_min_max: bool = True):
    """ExtractSignedDistanceField<float> on a device-resident occupancy grid [nx, ny, nz].

    Returns (sdf float32 [nx, ny, nz], min_max float32 [2] or None). Asynchronous."""
    _require_cuda(occupancy, torch.float32, "occupancy")
    if occupancy.dim() != 3:
        raise ValueError("occupancy must be indexed [x, y, z]")
    device = occupancy.device
    if out is None:
        out = torch.empty_like(occupancy)
    _require_cuda(out, torch.float32, "out")
    if compute_min_max and min_max is None:
        min_max = torch.empty(2, dtype=torch.float32, device=device)
    nx, ny, nz = occupancy.shape
    code = _capi.library().vgt_b200_sdf_f32_dev(
        occupancy.data_ptr(), nx, ny, nz, float(resolution), int(unknown_is_filled),
        int(add_virtual_border), device.index or 0, out.data_ptr(),
        min_max.data_ptr() if compute_min_max else None, _stream_handle(device))
    _capi.check(code)
    return out, (min_max if compute_min_max else None)


def signed_distance_field_profile(occupancy: torch.Tensor, resolution: float,
                                  out: torch.Tensor, min_max: torch.Tensor | None = None,
                                  kernels: bool = False):
    """Same kernels as signed_distance_field with CUDA events between them; synchronises.
    Returns [ms z-scan, ms y-pass, ms x-pass + finalize]; kernels=True appends the durations of
    the y pass's and the x pass's main window-kernel launch alone (a pass also holds the pilot
    probe, the decision and the stack kernel over the hand-over list)."""
    _require_cuda(occupancy, torch.float32, "occupancy")
    _require_cuda(out, torch.float32, "out")
    device = occupancy.device
    nx, ny, nz = occupancy.shape
    pass_ms = (ctypes.c_float * 5)()
    code = _capi.library().vgt_b200_sdf_f32_dev_profile(
        occupancy.data_ptr(), nx, ny, nz, float(resolution), 1, 0, device.index or 0,
        out.data_ptr(), None if min_max is None else min_max.data_ptr(), _stream_handle(device),
        pass_ms)
    _capi.check(code)
    return [float(v) for v in pass_ms][:5 if kernels else 3]


def signed_distance_field_f64(occupancy: torch.Tensor, resolution: float,
                              unknown_is_filled: bool = True, add_virtual_border: bool = False):
    _require_cuda(occupancy, torch.float32, "occupancy")
    device = occupancy.device
    out = torch.empty(occupancy.shape, dtype=torch.float64, device=device)
    min_max = torch.empty(2, dtype=torch.float64, device=device)
    nx, ny, nz = occupancy.shape
    code = _capi.library().vgt_b200_sdf_f64_dev(
        occupancy.data_ptr(), nx, ny, nz, float(resolution), int(unknown_is_filled),
        int(add_virtual_border), device.index or 0, out.data_ptr(), min_max.data_ptr(),
        _stream_handle(device))
    _capi.check(code)
    return out, min_max


def signed_distance_field_from_mask(mask: torch.Tensor, resolution: float,
                                    add_virtual_border: bool = False):
    _require_cuda(mask, torch.uint8, "mask")
    device = mask.device
    out = torch.empty(mask.shape, dtype=torch.float32, device=device)
    min_max = torch.empty(2, dtype=torch.float32, device=device)
    nx, ny, nz = mask.shape
    code = _capi.library().vgt_b200_sdf_from_mask_f32_dev(
        mask.data_ptr(), nx, ny, nz, float(resolution), int(add_virtual_border),
        device.index or 0, out.data_ptr(), min_max.data_ptr(), _stream_handle(device))
    _capi.check(code)
    return out, min_max


def edt_local_passes(occupancy_slab: torch.Tensor, unknown_is_filled: bool = True,
                     out: torch.Tensor | None = None, send_parts: int = 0) -> torch.Tensor:
    """z and y passes on an x-slab [nx_local, ny, nz] -> sign-fused int32 words.

    send_parts = G > 1 writes the result in send layout (G blocks [nx_local, rows_h, nz] back to
    back, block h = the part of every y-line that rank h owns); raises NotImplementedError when the
    library cannot (ny > 1024), in which case the caller packs."""
    _require_cuda(occupancy_slab, torch.float32, "occupancy_slab")
    device = occupancy_slab.device
    if out is None:
        out = torch.empty(occupancy_slab.shape, dtype=torch.int32, device=device)
    _require_cuda(out, torch.int32, "out")
    nx, ny, nz = occupancy_slab.shape
    code = _capi.library().vgt_b200_edt_local_passes_dev(
        occupancy_slab.data_ptr(), nx, ny, nz, int(unknown_is_filled), int(send_parts),
        device.index or 0, out.data_ptr(), _stream_handle(device))
    if code == _capi.ERR_UNSUPPORTED:
        raise NotImplementedError(_capi.last_error())
    _capi.check(code)
    return out


def edt_local_passes_scatter(occupancy_slab: torch.Tensor, rank: int, x_offset: int,
                             nx_total: int, peer_buffer_ptrs, receive_capacity_words: int,
                             unknown_is_filled: bool = True) -> None:
    """z and y passes on an x-slab with the exchange fused in: the y pass stores each rank's
    part of every line straight into that rank's receive buffer (peer-mapped device pointers,
    entry h = rank h's buffer laid out [nx_total, rows_h, nz], every buffer
    ``receive_capacity_words`` int32 words long; the library refuses a layout that does not fit)."""
    _require_cuda(occupancy_slab, torch.float32, "occupancy_slab")
    device = occupancy_slab.device
    nx, ny, nz = occupancy_slab.shape
    pointers = (ctypes.c_uint64 * len(peer_buffer_ptrs))(*[int(p) for p in peer_buffer_ptrs])
    code = _capi.library().vgt_b200_edt_local_passes_scatter_dev(
        occupancy_slab.data_ptr(), nx, ny, nz, int(unknown_is_filled), len(peer_buffer_ptrs),
        int(rank), int(x_offset), int(nx_total), pointers, int(receive_capacity_words),
        device.index or 0, _stream_handle(device))
    if code == _capi.ERR_UNSUPPORTED:
        raise NotImplementedError(_capi.last_error())
    _capi.check(code)


def edt_final_pass(packed: torch.Tensor, y_offset: int, ny_total: int, resolution: float,
                   add_virtual_border: bool = False, compute_min_max: bool = True):
    """x pass + finalize on a y-slab [nx, ny_local, nz] of sign-fused words. ``packed`` is
    destroyed (the envelope stacks are built in place in it); the SDF is a new tensor."""
    _require_cuda(packed, torch.int32, "packed")
    device = packed.device
    nx, ny_local, nz = packed.shape
    out = torch.empty(packed.shape, dtype=torch.float32, device=device)
    min_max = torch.empty(2, dtype=torch.float32, device=device) if compute_min_max else None
    code = _capi.library().vgt_b200_edt_final_pass_f32_dev(
        packed.data_ptr(), nx, ny_local, nz, int(y_offset), int(ny_total), float(resolution),
        int(add_virtual_border), device.index or 0, out.data_ptr(),
        min_max.data_ptr() if compute_min_max else None, _stream_handle(device))
    _capi.check(code)
    return out, min_max


def raycast_cloud(points_xyz: torch.Tensor, x_gc, max_range: float, counts: torch.Tensor,
                  voxel_size: float) -> torch.Tensor:
    """Accumulates one cloud (float64 or float32 [N, 3], cloud frame) into counts int32
    [nx, ny, nz, 2]; float32 points are widened on the device."""
    single = points_xyz.dtype == torch.float32
    _require_cuda(points_xyz, torch.float32 if single else torch.float64, "points_xyz")
    _require_cuda(counts, torch.int32, "counts")
    if counts.dim() != 4 or counts.shape[3] != 2:
        raise ValueError("counts must be [nx, ny, nz, 2]")
    device = counts.device
    column_major = np.ascontiguousarray(np.asarray(x_gc, dtype=np.float64).reshape(4, 4).T)
    nx, ny, nz, _ = counts.shape
    entry = _capi.library().vgt_b200_raycast_f32_dev if single \
        else _capi.library().vgt_b200_raycast_f64_dev
    code = entry(
        points_xyz.data_ptr(), points_xyz.shape[0],
        column_major.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), float(max_range), nx, ny,
        nz, float(voxel_size), device.index or 0, counts.data_ptr(), _stream_handle(device))
    _capi.check(code)
    return counts


def filter_grids(counts: torch.Tensor, occupancy: torch.Tensor,
                 filter_options: PointCloudVoxelizationFilterOptions) -> torch.Tensor:
    """counts int32 [num_grids, nx, ny, nz, 2]; occupancy float32 [nx, ny, nz] updated in place."""
    _require_cuda(counts, torch.int32, "counts")
    _require_cuda(occupancy, torch.float32, "occupancy")
    device = occupancy.device
    options = filter_options.as_struct()
    code = _capi.library().vgt_b200_filter_dev(
        counts.data_ptr(), counts.shape[0], occupancy.numel(), ctypes.byref(options),
        device.index or 0, occupancy.data_ptr(), _stream_handle(device))
    _capi.check(code)
    return occupancy


# --------------------------------------------------------------------------------------------------
# SignedDistanceField queries on a device-resident SDF
# --------------------------------------------------------------------------------------------------
QUERY_NO_VALUE, QUERY_VALUE, QUERY_THROWS = 0, 1, 2


class DeviceSignedDistanceField:
    """A SignedDistanceField<float> that stays on the GPU: the float grid [nx, ny, nz] plus
    resolution and origin transform. The batched queries mirror the reference's per-point
    members (signed_distance_field.hpp:808-1203); points are float64 [N, 3] CUDA tensors in the
    WORLD frame. Every query returns (values, valid uint8): 1 = value, 0 = the reference returns
    an empty query, 2 = the reference throws for that point."""

    def __init__(self, sdf: torch.Tensor, resolution: float, origin_transform=None):
        _require_cuda(sdf, torch.float32, "sdf")
        if sdf.dim() != 3:
            raise ValueError("sdf must be indexed [x, y, z]")
        self.sdf = sdf
        self.resolution = float(resolution)
        self.origin_transform = np.eye(4) if origin_transform is None else \
            np.asarray(origin_transform, dtype=np.float64).reshape(4, 4)
        self._view = _capi.SdfView()
        self._view.d_sdf = sdf.data_ptr()
        self._view.nx, self._view.ny, self._view.nz = (int(d) for d in sdf.shape)
        self._view.resolution = self.resolution
        column_major = np.ascontiguousarray(self.origin_transform.T).reshape(-1)
        for i in range(16):
            self._view.origin_transform[i] = float(column_major[i])

    def _points(self, points: torch.Tensor) -> torch.Tensor:
        _require_cuda(points, torch.float64, "points")
        if points.dim() != 2 or points.shape[1] != 3:
            raise ValueError("points must be [N, 3]")
        if points.device != self.sdf.device:
            raise ValueError("points and sdf must live on the same device")
        return points

    def _launch(self, name, points, width, *middle):
        points = self._points(points)
        device = self.sdf.device
        count = points.shape[0]
        shape = (count,) if width == 1 else (count, width)
        values = torch.empty(shape, dtype=torch.float64, device=device)
        valid = torch.empty(count, dtype=torch.uint8, device=device)
        function = getattr(_capi.library(), name)
        code = function(ctypes.byref(self._view), points.data_ptr(), count, *middle,
                        device.index or 0, values.data_ptr(), valid.data_ptr(),
                        _stream_handle(device))
        _capi.check(code)
        return values, valid

    def EstimateLocationDistance(self, points: torch.Tensor):
        return self._launch("vgt_b200_sdf_estimate_distance_dev", points, 1)

    def GetLocationCoarseGradient(self, points: torch.Tensor, enable_edge_gradients=False):
        return self._launch("vgt_b200_sdf_coarse_gradient_dev", points, 3,
                            int(bool(enable_edge_gradients)))

    def GetLocationFineGradient(self, points: torch.Tensor, nominal_window_size: float):
        return self._launch("vgt_b200_sdf_fine_gradient_dev", points, 3,
                            float(nominal_window_size))

    def ProjectLocationOutOfCollisionToMinimumDistance(
            self, points: torch.Tensor, minimum_distance: float = 0.0,
            stepsize_multiplier: float = 1.0 / 10.0, max_steps: int = 1_000_000):
        return self._launch("vgt_b200_sdf_project_out_of_collision_dev", points, 3,
                            float(minimum_distance), float(stepsize_multiplier), int(max_steps))

    def ProjectLocationOutOfCollision(self, points: torch.Tensor,
                                      stepsize_multiplier: float = 1.0 / 10.0):
        return self.ProjectLocationOutOfCollisionToMinimumDistance(points, 0.0,
                                                                   stepsize_multiplier)

    def ComputeLocalExtremaMap(self) -> torch.Tensor:
        """float64 [nx, ny, nz, 3]: per cell the grid-frame centre of the cell its gradient walk
        ends at, +inf when it leaves the grid (signed_distance_field.hpp:1207-1231)."""
        device = self.sdf.device
        extrema = torch.empty(tuple(self.sdf.shape) + (3,), dtype=torch.float64, device=device)
        code = _capi.library().vgt_b200_sdf_local_extrema_map_dev(
            ctypes.byref(self._view), device.index or 0, extrema.data_ptr(),
            _stream_handle(device))
        _capi.check(code)
        return extrema
