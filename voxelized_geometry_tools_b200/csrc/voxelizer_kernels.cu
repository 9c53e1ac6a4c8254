// sm_100a raycasting voxelizer + filter and their C-ABI entry points.
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: the DDA must reproduce the double-precision
// CPU path (cpu_pointcloud_voxelization.cpp:208-436) operation for operation, and a contracted
// multiply-add would change which voxel boundary a ray crosses first.
//
// Data layout: tracking grid = int32[voxel][2] = {seen_free, seen_filled}, voxel index
// x*(ny*nz) + y*nz + z -- the CpuVoxelizationTrackingCell layout (cpu_pcv.hpp:24-32).
#include <cmath>
#include <chrono>
#include <limits>
#include <vector>

#include "common.cuh"
#include "host_transfer.cuh"

namespace vgt_b200
{
namespace voxelizer
{
namespace
{
struct GridFrame
{
  int64_t nx;
  int64_t ny;
  int64_t nz;
  double voxel_size;
  double inverse_voxel_size;
  double extent_x;
  double extent_y;
  double extent_z;
};

struct CloudPose
{
  double m[16];  // X_GC, column-major
};

struct Cell
{
  int64_t x;
  int64_t y;
  int64_t z;
};

__device__ __forceinline__ bool InGrid(const GridFrame& g, const Cell& c)
{
  return c.x >= 0 && c.x < g.nx && c.y >= 0 && c.y < g.ny && c.z >= 0 && c.z < g.nz;
}

__device__ __forceinline__ Cell CellOf(const GridFrame& g, double px, double py, double pz)
{
  return Cell{static_cast<int64_t>(floor(px * g.inverse_voxel_size)),
              static_cast<int64_t>(floor(py * g.inverse_voxel_size)),
              static_cast<int64_t>(floor(pz * g.inverse_voxel_size))};
}

__device__ __forceinline__ void Bump(const GridFrame& g, int32_t* counts, const Cell& c, int which)
{
  const int64_t voxel = (c.x * g.ny + c.y) * g.nz + c.z;
  atomicAdd(counts + 2 * voxel + which, 1);  // result unused -> RED.ADD
}

__device__ __forceinline__ int StepToward(int64_t difference)
{
  return (difference > 0) ? 1 : ((difference < 0) ? -1 : 0);
}

// cpu_pcv.cpp:336-354.
__device__ __forceinline__ double FirstBoundaryT(double point_axis, double ray_axis,
                                                 double cell_low, double cell_high)
{
  if (ray_axis > 0.0)
  {
    return fabs((cell_high - point_axis) / ray_axis);
  }
  else if (ray_axis < -0.0)
  {
    return fabs((point_axis - cell_low) / ray_axis);
  }
  return __longlong_as_double(0x7ff0000000000000ll);  // +inf
}

// One thread per point: the literal double-precision DDA of
// CpuPointCloudVoxelizer::DoRaycastSinglePoint (cpu_pcv.cpp:208-436).
// Scalar = double, or float for clouds that arrive as float32 (the reference's ROS wrapper,
// pointcloud_voxelization_ros_interface.hpp:35-97, hands PointCloud2 floats over through
// CopyPointLocationIntoDoublePtr, i.e. widened to double exactly): 12 instead of 24 bytes per
// point over the bus, the widening happens here.
template <typename Scalar>
__global__ void __launch_bounds__(128) RaycastCloudKernel(
    const Scalar* __restrict__ points, int64_t num_points, CloudPose pose, double max_range,
    GridFrame grid, int32_t* counts)
{
  const int64_t index = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (index >= num_points)
  {
    return;
  }
  const double cx = static_cast<double>(points[3 * index + 0]);
  const double cy = static_cast<double>(points[3 * index + 1]);
  const double cz = static_cast<double>(points[3 * index + 2]);
  // Skip NaN / infinite points (cpu_pcv.cpp:191-192).
  if (!(isfinite(cx) && isfinite(cy) && isfinite(cz)))
  {
    return;
  }
  const double* m = pose.m;
  // p_GP = X_GC * p_CP, summed left to right per row (cpu_pcv.cpp:195).
  const double gx = ((m[0] * cx + m[4] * cy) + m[8] * cz) + m[12];
  const double gy = ((m[1] * cx + m[5] * cy) + m[9] * cz) + m[13];
  const double gz = ((m[2] * cx + m[6] * cy) + m[10] * cz) + m[14];
  const double ox = m[12];
  const double oy = m[13];
  const double oz = m[14];

  // Step 1 (cpu_pcv.cpp:217-226).
  const double rx = gx - ox;
  const double ry = gy - oy;
  const double rz = gz - oz;
  const double ray_length = sqrt((rx * rx + ry * ry) + rz * rz);
  const bool clipped = ray_length > max_range;
  double fx = gx;
  double fy = gy;
  double fz = gz;
  if (clipped)
  {
    const double scale = max_range / ray_length;
    fx = ox + rx * scale;
    fy = oy + ry * scale;
    fz = oz + rz * scale;
  }

  // Step 2 (cpu_pcv.cpp:229-290).
  double sx = ox;
  double sy = oy;
  double sz = oz;
  const Cell origin_cell = CellOf(grid, ox, oy, oz);
  if (!InGrid(grid, origin_cell))
  {
    // A point on the sensor origin (e.g. a zero-filled invalid depth return): the direction is
    // 0 / 0, the start point NaN. On the CPU the NaN start index converts to INT64_MIN, is out
    // of bounds and the ray marks nothing (its final voxel, the origin's, is outside the grid
    // too); a device float -> int conversion of NaN gives 0 instead and would walk from voxel
    // (0, 0, 0). Same outcome as the CPU path, stated explicitly.
    if (!(ray_length > 0.0))
    {
      return;
    }
    double t_enter = 0.0;
    double t_exit = max_range;
    const double direction[3] = {rx / ray_length, ry / ray_length, rz / ray_length};
    const double origin[3] = {ox, oy, oz};
    const double extent[3] = {grid.extent_x, grid.extent_y, grid.extent_z};
    const double flat_threshold = 1e-10;
#pragma unroll
    for (int axis = 0; axis < 3; axis++)
    {
      if (fabs(direction[axis]) < flat_threshold)
      {
        const bool inside_slab = origin[axis] >= 0.0 && origin[axis] < extent[axis];
        if (!inside_slab)
        {
          return;
        }
      }
      else
      {
        const double inverse = 1.0 / direction[axis];
        const double t_low = (0.0 - origin[axis]) * inverse;
        const double t_high = (extent[axis] - origin[axis]) * inverse;
        const double t_near = (t_low <= t_high) ? t_low : t_high;
        const double t_far = (t_low <= t_high) ? t_high : t_low;
        if (t_near > t_enter)
        {
          t_enter = t_near;
        }
        // cpu_pcv.cpp:274-277: the exit bound only ever grows; mirrored on purpose.
        if (t_far > t_exit)
        {
          t_exit = t_far;
        }
        if (t_enter > t_exit)
        {
          return;
        }
      }
    }
    const double advance = t_enter + 1e-10;
    sx = ox + direction[0] * advance;
    sy = oy + direction[1] * advance;
    sz = oz + direction[2] * advance;
  }

  // Steps 3-5 (cpu_pcv.cpp:293-365).
  const Cell start_cell = CellOf(grid, sx, sy, sz);
  const Cell final_cell = CellOf(grid, fx, fy, fz);
  const int64_t step_x = StepToward(final_cell.x - start_cell.x);
  const int64_t step_y = StepToward(final_cell.y - start_cell.y);
  const int64_t step_z = StepToward(final_cell.z - start_cell.z);
  const double half = grid.voxel_size * 0.5;
  const double centre_x = grid.voxel_size * (static_cast<double>(start_cell.x) + 0.5);
  const double centre_y = grid.voxel_size * (static_cast<double>(start_cell.y) + 0.5);
  const double centre_z = grid.voxel_size * (static_cast<double>(start_cell.z) + 0.5);
  double tx = FirstBoundaryT(sx, rx, centre_x - half, centre_x + half);
  double ty = FirstBoundaryT(sy, ry, centre_y - half, centre_y + half);
  double tz = FirstBoundaryT(sz, rz, centre_z - half, centre_z + half);
  const double dtx = fabs(grid.voxel_size / rx);
  const double dty = fabs(grid.voxel_size / ry);
  const double dtz = fabs(grid.voxel_size / rz);

  // Step 6 (cpu_pcv.cpp:368-381).
  if (InGrid(grid, final_cell))
  {
    Bump(grid, counts, final_cell, clipped ? 0 : 1);
  }

  // Walk (cpu_pcv.cpp:384-435).
  Cell at = start_cell;
  while (at.x != final_cell.x || at.y != final_cell.y || at.z != final_cell.z)
  {
    if (!InGrid(grid, at))
    {
      break;
    }
    Bump(grid, counts, at, 0);
    if (tx <= ty && tx <= tz)
    {
      if (at.x == final_cell.x)
      {
        break;
      }
      at.x += step_x;
      tx += dtx;
    }
    else if (ty <= tx && ty <= tz)
    {
      if (at.y == final_cell.y)
      {
        break;
      }
      at.y += step_y;
      ty += dty;
    }
    else
    {
      if (at.z == final_cell.z)
      {
        break;
      }
      at.z += step_z;
      tz += dtz;
    }
  }
}

// One thread per voxel: per-camera rule (pcv_if.hpp:55-86) + combine (cpu_pcv.cpp:448-490).
__global__ void __launch_bounds__(256) FilterGridsKernel(
    const int2* __restrict__ counts, int32_t num_grids, int64_t num_voxels,
    double percent_seen_free, int32_t outlier_points_threshold, int32_t num_cameras_seen_free,
    float* occupancy)
{
  const int64_t voxel = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (voxel >= num_voxels)
  {
    return;
  }
  const float current = occupancy[voxel];
  if (!(static_cast<double>(current) <= 0.5))
  {
    return;  // filled cells stay filled
  }
  int32_t cameras_free = 0;
  int32_t cameras_filled = 0;
  for (int32_t g = 0; g < num_grids; g++)
  {
    const int2 cell = __ldcs(counts + static_cast<int64_t>(g) * num_voxels + voxel);
    const int32_t seen_free = cell.x;
    const int32_t seen_filled = (cell.y >= outlier_points_threshold) ? cell.y : 0;
    if (seen_free > 0 && seen_filled > 0)
    {
      const double fraction_free =
          static_cast<double>(seen_free) / static_cast<double>(seen_free + seen_filled);
      if (fraction_free >= percent_seen_free)
      {
        cameras_free += 1;
      }
      else
      {
        cameras_filled += 1;
      }
    }
    else if (seen_free > 0)
    {
      cameras_free += 1;
    }
    else if (seen_filled > 0)
    {
      cameras_filled += 1;
    }
  }
  float result;
  if (cameras_filled > 0)
  {
    result = 1.0f;
  }
  else if (cameras_free >= num_cameras_seen_free)
  {
    result = 0.0f;
  }
  else
  {
    result = 0.5f;
  }
  occupancy[voxel] = result;
}

int CheckFilter(const vgt_b200_filter_options* filter)
{
  if (filter == nullptr)
  {
    return FailInvalid("null filter options");
  }
  // Same conditions as the PointCloudVoxelizationFilterOptions constructor (pcv_if.hpp:30-41).
  if (!(filter->percent_seen_free > 0.0) || filter->percent_seen_free > 1.0)
  {
    return FailInvalid("0 < percent_seen_free_ <= 1 must be true");
  }
  if (filter->outlier_points_threshold <= 0)
  {
    return FailInvalid("outlier_points_threshold_ <= 0");
  }
  if (filter->num_cameras_seen_free <= 0)
  {
    return FailInvalid("num_cameras_seen_free_ <= 0");
  }
  return VGT_B200_OK;
}

GridFrame MakeFrame(int64_t nx, int64_t ny, int64_t nz, double voxel_size)
{
  GridFrame g;
  g.nx = nx;
  g.ny = ny;
  g.nz = nz;
  g.voxel_size = voxel_size;
  g.inverse_voxel_size = 1.0 / voxel_size;
  g.extent_x = static_cast<double>(nx) * voxel_size;
  g.extent_y = static_cast<double>(ny) * voxel_size;
  g.extent_z = static_cast<double>(nz) * voxel_size;
  return g;
}

template <typename Scalar>
int LaunchRaycast(const Scalar* d_points, int64_t num_points, const double* x_gc,
                  double max_range, const GridFrame& grid, int32_t* d_counts, cudaStream_t stream)
{
  if (num_points <= 0)
  {
    return VGT_B200_OK;
  }
  CloudPose pose;
  for (int i = 0; i < 16; i++)
  {
    pose.m[i] = x_gc[i];
  }
  const int threads = 128;
  const int64_t blocks = (num_points + threads - 1) / threads;
  RaycastCloudKernel<Scalar><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
      d_points, num_points, pose, max_range, grid, d_counts); NoteKernelLaunch();
  VGT_CUDA_TRY(cudaGetLastError(), "RaycastCloudKernel launch");
  return VGT_B200_OK;
}

int LaunchFilter(const int32_t* d_counts, int32_t num_grids, int64_t num_voxels,
                 const vgt_b200_filter_options& filter, float* d_occupancy, cudaStream_t stream)
{
  const int threads = 256;
  const int64_t blocks = (num_voxels + threads - 1) / threads;
  FilterGridsKernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
      reinterpret_cast<const int2*>(d_counts), num_grids, num_voxels, filter.percent_seen_free,
      filter.outlier_points_threshold, filter.num_cameras_seen_free, d_occupancy); NoteKernelLaunch();
  VGT_CUDA_TRY(cudaGetLastError(), "FilterGridsKernel launch");
  return VGT_B200_OK;
}

int CheckGrid(int64_t nx, int64_t ny, int64_t nz, double voxel_size)
{
  if (!ValidDims(nx, ny, nz))
  {
    return FailInvalid("grid dimensions out of range");
  }
  if (!(voxel_size > 0.0) || !std::isfinite(voxel_size))
  {
    return FailInvalid("voxel_size must be positive and finite");
  }
  return VGT_B200_OK;
}
template <typename Scalar>
struct CloudOf;
template <>
struct CloudOf<double>
{
  using Type = vgt_b200_cloud;
};
template <>
struct CloudOf<float>
{
  using Type = vgt_b200_cloud_f32;
};

struct OwnedStream
{
  cudaStream_t stream = nullptr;
  ~OwnedStream()
  {
    if (stream != nullptr)
    {
      // (nothing of this call may still be queued when its buffers go back to the pool)
      cudaStreamSynchronize(stream);
      cudaStreamDestroy(stream);
    }
  }
};

struct OwnedEvent
{
  cudaEvent_t event = nullptr;
  ~OwnedEvent()
  {
    if (event != nullptr)
    {
      cudaEventDestroy(event);
    }
  }
};

// <Backend>PointCloudVoxelizer::DoVoxelizePointClouds on host buffers. Two streams: the static
// map travels on the copy stream while the clouds are raycast on the compute stream; the points
// of cloud c + 1 are uploaded (into the other of two buffers) while cloud c is raycast; pageable
// buffers go through the pinned staging ring (host_transfer.cuh). The two durations of
// VoxelizerRuntime come from CUDA events, so the host never waits in the middle of the call.
template <typename Scalar>
int VoxelizeFromHost(
    const float* static_occupancy, int64_t nx, int64_t ny, int64_t nz, double voxel_size,
    const typename CloudOf<Scalar>::Type* clouds, int32_t num_clouds,
    const vgt_b200_filter_options* filter, int device, float* out_occupancy, int32_t* out_counts,
    double* out_seconds)
{
  int check = CheckGrid(nx, ny, nz, voxel_size);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  check = CheckFilter(filter);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  if (static_occupancy == nullptr || out_occupancy == nullptr || num_clouds < 0
      || (num_clouds > 0 && clouds == nullptr))
  {
    return FailInvalid("null pointer or negative cloud count");
  }
  for (int32_t c = 0; c < num_clouds; c++)
  {
    if (clouds[c].num_points < 0 || (clouds[c].num_points > 0 && clouds[c].points_xyz == nullptr))
    {
      // pcv_if.hpp:281-289 rejects null clouds with invalid_argument.
      return FailInvalid("pointclouds[%d] is null", c);
    }
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  KeepPoolMemory(device);
  const int64_t num_voxels = nx * ny * nz;
  const int64_t num_grids = num_clouds > 0 ? num_clouds : 1;
  const GridFrame grid = MakeFrame(nx, ny, nz, voxel_size);
  OwnedStream compute, copy;
  VGT_CUDA_TRY(cudaStreamCreateWithFlags(&compute.stream, cudaStreamNonBlocking), "stream");
  VGT_CUDA_TRY(cudaStreamCreateWithFlags(&copy.stream, cudaStreamNonBlocking), "stream");
  OwnedEvent started, raycast_done, filter_done, map_arrived, allocated;
  VGT_CUDA_TRY(cudaEventCreate(&started.event), "event");
  VGT_CUDA_TRY(cudaEventCreate(&raycast_done.event), "event");
  VGT_CUDA_TRY(cudaEventCreate(&filter_done.event), "event");
  VGT_CUDA_TRY(cudaEventCreateWithFlags(&map_arrived.event, cudaEventDisableTiming), "event");
  VGT_CUDA_TRY(cudaEventCreateWithFlags(&allocated.event, cudaEventDisableTiming), "event");
  OwnedEvent buffer_free[2];
  vgt_b200::StagedTransfer points_transfer;
  vgt_b200::StagedTransfer map_transfer;

  StreamScratch<int32_t> d_counts;
  StreamScratch<float> d_occupancy;
  StreamScratch<Scalar> d_points[2];
  VGT_CUDA_TRY(d_counts.Allocate(2 * num_voxels * num_grids, compute.stream),
               "tracking grid allocation");
  VGT_CUDA_TRY(d_occupancy.Allocate(num_voxels, compute.stream), "occupancy allocation");
  int64_t max_points = 0;
  for (int32_t c = 0; c < num_clouds; c++)
  {
    max_points = (clouds[c].num_points > max_points) ? clouds[c].num_points : max_points;
  }
  if (max_points > 0)
  {
    VGT_CUDA_TRY(d_points[0].Allocate(3 * max_points, compute.stream), "points allocation");
    VGT_CUDA_TRY(d_points[1].Allocate(3 * max_points, compute.stream), "points allocation");
  }
  // Declared after the buffers and the staging objects, so it runs first on every exit path:
  // nothing of this call is still queued when they go back to the pool / the slot cache.
  struct DrainOnExit
  {
    cudaStream_t a;
    cudaStream_t b;
    ~DrainOnExit()
    {
      cudaStreamSynchronize(a);
      cudaStreamSynchronize(b);
    }
  } drain{compute.stream, copy.stream};
  VGT_CUDA_TRY(cudaEventRecord(allocated.event, compute.stream), "event record");
  VGT_CUDA_TRY(cudaStreamWaitEvent(copy.stream, allocated.event, 0), "stream wait");
  VGT_CUDA_TRY(cudaEventRecord(started.event, compute.stream), "event record");
  VGT_CUDA_TRY(cudaMemsetAsync(d_counts.get(), 0, sizeof(int32_t) * 2 * num_voxels * num_grids,
                               compute.stream),
               "zero tracking grids");
  // the static map, on its own stream: only the filter needs it
  {
    const size_t bytes = sizeof(float) * static_cast<size_t>(num_voxels);
    VGT_CUDA_TRY(map_transfer.ToDevice(reinterpret_cast<char*>(d_occupancy.get()), bytes,
                                       reinterpret_cast<const char*>(static_occupancy), bytes,
                                       bytes, 1, copy.stream),
                 "copy occupancy to device");
    VGT_CUDA_TRY(cudaEventRecord(map_arrived.event, copy.stream), "event record");
  }
  // clouds: upload into buffer c % 2 on the copy stream, raycast on the compute stream
  OwnedEvent uploaded[2];
  for (int b = 0; b < 2; b++)
  {
    VGT_CUDA_TRY(cudaEventCreateWithFlags(&uploaded[b].event, cudaEventDisableTiming), "event");
    VGT_CUDA_TRY(cudaEventCreateWithFlags(&buffer_free[b].event, cudaEventDisableTiming),
                 "event");
  }
  int used = 0;
  for (int32_t c = 0; c < num_clouds; c++)
  {
    if (clouds[c].num_points == 0)
    {
      continue;
    }
    const int b = used & 1;
    if (used >= 2)
    {
      // the raycast that last read this buffer must have finished
      VGT_CUDA_TRY(cudaStreamWaitEvent(copy.stream, buffer_free[b].event, 0), "stream wait");
    }
    const size_t bytes = sizeof(Scalar) * 3 * static_cast<size_t>(clouds[c].num_points);
    VGT_CUDA_TRY(points_transfer.ToDevice(reinterpret_cast<char*>(d_points[b].get()), bytes,
                                          reinterpret_cast<const char*>(clouds[c].points_xyz),
                                          bytes, bytes, 1, copy.stream),
                 "copy points to device");
    VGT_CUDA_TRY(cudaEventRecord(uploaded[b].event, copy.stream), "event record");
    VGT_CUDA_TRY(cudaStreamWaitEvent(compute.stream, uploaded[b].event, 0), "stream wait");
    const int status = LaunchRaycast<Scalar>(d_points[b].get(), clouds[c].num_points,
                                             clouds[c].x_gc, clouds[c].max_range, grid,
                                             d_counts.get() + 2 * num_voxels * c, compute.stream);
    if (status != VGT_B200_OK)
    {
      return status;
    }
    VGT_CUDA_TRY(cudaEventRecord(buffer_free[b].event, compute.stream), "event record");
    used++;
  }
  VGT_CUDA_TRY(cudaEventRecord(raycast_done.event, compute.stream), "event record");
  VGT_CUDA_TRY(cudaStreamWaitEvent(compute.stream, map_arrived.event, 0), "stream wait");
  const int status = LaunchFilter(d_counts.get(), num_clouds, num_voxels, *filter,
                                  d_occupancy.get(), compute.stream);
  if (status != VGT_B200_OK)
  {
    return status;
  }
  VGT_CUDA_TRY(cudaEventRecord(filter_done.event, compute.stream), "event record");
  {
    const size_t bytes = sizeof(float) * static_cast<size_t>(num_voxels);
    VGT_CUDA_TRY(map_transfer.ToHost(reinterpret_cast<char*>(out_occupancy), bytes,
                                     reinterpret_cast<const char*>(d_occupancy.get()), bytes,
                                     bytes, 1, compute.stream),
                 "copy occupancy to host");
  }
  if (out_counts != nullptr && num_clouds > 0)
  {
    const size_t bytes = sizeof(int32_t) * 2 * static_cast<size_t>(num_voxels * num_clouds);
    VGT_CUDA_TRY(map_transfer.ToHost(reinterpret_cast<char*>(out_counts), bytes,
                                     reinterpret_cast<const char*>(d_counts.get()), bytes, bytes,
                                     1, compute.stream),
                 "copy counts to host");
  }
  VGT_CUDA_TRY(cudaStreamSynchronize(copy.stream), "uploads");
  VGT_CUDA_TRY(cudaStreamSynchronize(compute.stream), "voxelization");
  if (out_seconds != nullptr)
  {
    float raycast_ms = 0.0f;
    float filter_ms = 0.0f;
    cudaEventElapsedTime(&raycast_ms, started.event, raycast_done.event);
    cudaEventElapsedTime(&filter_ms, raycast_done.event, filter_done.event);
    out_seconds[0] = static_cast<double>(raycast_ms) * 1e-3;
    out_seconds[1] = static_cast<double>(filter_ms) * 1e-3;
  }
  return VGT_B200_OK;
}
}  // namespace
}  // namespace voxelizer
}  // namespace vgt_b200

using namespace vgt_b200;
using namespace vgt_b200::voxelizer;

extern "C"
{
int vgt_b200_raycast_f64_dev(
    const double* d_points_xyz, int64_t num_points, const double* x_gc, double max_range,
    int64_t nx, int64_t ny, int64_t nz, double voxel_size, int device, int32_t* d_counts,
    void* stream)
{
  const int check = CheckGrid(nx, ny, nz, voxel_size);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  if (x_gc == nullptr || d_counts == nullptr || (num_points > 0 && d_points_xyz == nullptr)
      || num_points < 0)
  {
    return FailInvalid("null pointer or negative point count");
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  return LaunchRaycast(d_points_xyz, num_points, x_gc, max_range,
                       MakeFrame(nx, ny, nz, voxel_size), d_counts,
                       static_cast<cudaStream_t>(stream));
}

int vgt_b200_filter_dev(
    const int32_t* d_counts, int32_t num_grids, int64_t num_voxels,
    const vgt_b200_filter_options* filter, int device, float* d_occupancy, void* stream)
{
  const int check = CheckFilter(filter);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  if (num_grids < 0 || num_voxels < 1 || d_occupancy == nullptr
      || (num_grids > 0 && d_counts == nullptr))
  {
    return FailInvalid("bad grid count or null pointer");
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  return LaunchFilter(d_counts, num_grids, num_voxels, *filter, d_occupancy,
                      static_cast<cudaStream_t>(stream));
}

int vgt_b200_raycast_f32_dev(
    const float* d_points_xyz, int64_t num_points, const double* x_gc, double max_range,
    int64_t nx, int64_t ny, int64_t nz, double voxel_size, int device, int32_t* d_counts,
    void* stream)
{
  const int check = CheckGrid(nx, ny, nz, voxel_size);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  if (x_gc == nullptr || d_counts == nullptr || (num_points > 0 && d_points_xyz == nullptr)
      || num_points < 0)
  {
    return FailInvalid("null pointer or negative point count");
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  return LaunchRaycast(d_points_xyz, num_points, x_gc, max_range,
                       MakeFrame(nx, ny, nz, voxel_size), d_counts,
                       static_cast<cudaStream_t>(stream));
}

int vgt_b200_voxelize_f64(
    const float* static_occupancy, int64_t nx, int64_t ny, int64_t nz, double voxel_size,
    const vgt_b200_cloud* clouds, int32_t num_clouds, const vgt_b200_filter_options* filter,
    int device, float* out_occupancy, int32_t* out_counts, double* out_seconds)
{
  return VoxelizeFromHost<double>(static_occupancy, nx, ny, nz, voxel_size, clouds, num_clouds,
                                  filter, device, out_occupancy, out_counts, out_seconds);
}

int vgt_b200_voxelize_f32(
    const float* static_occupancy, int64_t nx, int64_t ny, int64_t nz, double voxel_size,
    const vgt_b200_cloud_f32* clouds, int32_t num_clouds, const vgt_b200_filter_options* filter,
    int device, float* out_occupancy, int32_t* out_counts, double* out_seconds)
{
  return VoxelizeFromHost<float>(static_occupancy, nx, ny, nz, voxel_size, clouds, num_clouds,
                                 filter, device, out_occupancy, out_counts, out_seconds);
}
}  // extern "C"
