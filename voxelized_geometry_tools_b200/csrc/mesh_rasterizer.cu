// sm_100a triangle-mesh rasterizer and its C-ABI entry points.
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false: whether a cell is "touched" is a
// double-precision comparison (mesh_rasterizer.cpp:158-182) and must come out as on the CPU,
// operation for operation (dot = (ax*bx + ay*by) + az*bz, cross by components, no contraction).
//
// Work decomposition: blockIdx.x = triangle, blockIdx.y = slice of that triangle's index box;
// the threads of a block walk the box cell by cell (linear index, z fastest, so neighbouring
// lanes write neighbouring cells). A touched cell's occupancy is set to 1.0f: a plain store,
// every writer writes the same value. The reference is one thread per triangle with the same
// stores on relaxed atomics (mesh_rasterizer.cpp:184-197, 205-230).
#include <cmath>

#include "common.cuh"
#include "host_transfer.cuh"

namespace vgt_b200
{
namespace rasterizer
{
namespace
{
struct V3
{
  double x, y, z;
};

__device__ __forceinline__ V3 Sub(const V3& a, const V3& b)
{
  return V3{a.x - b.x, a.y - b.y, a.z - b.z};
}
__device__ __forceinline__ V3 Add(const V3& a, const V3& b)
{
  return V3{a.x + b.x, a.y + b.y, a.z + b.z};
}
__device__ __forceinline__ V3 Scale(const V3& a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double Dot(const V3& a, const V3& b)
{
  return (a.x * b.x + a.y * b.y) + a.z * b.z;
}
__device__ __forceinline__ V3 Cross(const V3& a, const V3& b)
{
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// PointProjectsInsideTriangle's helper (mesh_rasterizer.cpp:30-38)
__device__ __forceinline__ bool SameSide(const V3& a, const V3& b, const V3& p1, const V3& p2)
{
  const V3 ab = Sub(b, a);
  return Dot(Cross(ab, Sub(p1, a)), Cross(ab, Sub(p2, a))) >= 0.0;
}

// ClosestPointOnLineSegment (mesh_rasterizer.cpp:46-58); ClampValue as min(1, max(0, ratio)) in
// std::min / std::max's comparison order, so a NaN ratio (zero-length edge) clamps to 0.
__device__ __forceinline__ V3 ClosestOnSegment(const V3& a, const V3& b, const V3& q)
{
  const V3 ab = Sub(b, a);
  const double ratio = Dot(ab, Sub(q, a)) / Dot(ab, ab);
  const double lower = (0.0 < ratio) ? ratio : 0.0;
  const double clamped = (lower < 1.0) ? lower : 1.0;
  return Add(a, Scale(ab, clamped));
}

// CalcClosestPointOnTriangle (mesh_rasterizer.cpp:60-103). The edge branch picks among the three
// edge points by their own squared norms, as the reference does.
__device__ __forceinline__ V3 ClosestOnTriangle(const V3& v1, const V3& v2, const V3& v3,
                                                const V3& normal, const V3& q)
{
  if (SameSide(v1, v2, v3, q) && SameSide(v2, v3, v1, q) && SameSide(v3, v1, v2, q))
  {
    const V3 v1q = Sub(q, v1);
    const double normal_squared = Dot(normal, normal);
    V3 projection{0.0, 0.0, 0.0};
    if (normal_squared > 0.0)
    {
      projection = Scale(normal, Dot(normal, v1q) / normal_squared);
    }
    return Add(v1, Sub(v1q, projection));
  }
  const V3 c12 = ClosestOnSegment(v1, v2, q);
  const V3 c23 = ClosestOnSegment(v2, v3, q);
  const V3 c31 = ClosestOnSegment(v3, v1, q);
  const double d12 = Dot(c12, c12);
  const double d23 = Dot(c23, c23);
  const double d31 = Dot(c31, c31);
  if (d12 <= d23 && d12 <= d31)
  {
    return c12;
  }
  if (d23 <= d12 && d23 <= d31)
  {
    return c23;
  }
  return c31;
}

struct Frame
{
  double x_wg[16];   // the map's origin transform, column-major
  double x_gw[16];   // its inverse
  int64_t nx, ny, nz;
  double resolution;
  double inverse_resolution;
  double max_check_radius_squared;
};

__device__ __forceinline__ V3 Transform(const double* m, const V3& p)
{
  return V3{((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12] * 1.0,
            ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13] * 1.0,
            ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14] * 1.0};
}

constexpr int kThreads = 128;
constexpr int kSlices = 4;
constexpr int kFlagNotContained = 1;
constexpr int kFlagBadIndex = 2;

// RasterizeTriangleImpl (mesh_rasterizer.cpp:105-203). `cells` = the map's raw cells, the float
// occupancy first in each `cell_words`-word cell (OccupancyCell: 1 word, occupancy_map.hpp:28-58;
// OccupancyComponentCell: 2 words, occupancy_component_map.hpp:29-70).
__global__ void __launch_bounds__(kThreads)
RasterizeTrianglesKernel(const double* __restrict__ vertices, int64_t num_vertices,
                         const int32_t* __restrict__ triangles, Frame frame, int enforce_contains,
                         float* __restrict__ cells, int cell_words, int* __restrict__ flags)
{
  const int64_t triangle = blockIdx.x;
  V3 v[3];
#pragma unroll
  for (int k = 0; k < 3; k++)
  {
    const int64_t index = triangles[3 * triangle + k];
    if (index < 0 || index >= num_vertices)
    {
      // vertices.at(...) throws std::out_of_range (mesh_rasterizer.cpp:122-125)
      if (threadIdx.x == 0)
      {
        atomicOr(flags, kFlagBadIndex);
      }
      return;
    }
    v[k] = V3{vertices[3 * index], vertices[3 * index + 1], vertices[3 * index + 2]};
  }
  const V3 normal = Cross(Sub(v[1], v[0]), Sub(v[2], v[0]));
  // std::min({a, b, c}) / std::max({a, b, c}) in their comparison order (mesh_rasterizer.cpp:137-143)
  const auto min3 = [](double a, double b, double c)
  {
    double result = a;
    if (b < result) { result = b; }
    if (c < result) { result = c; }
    return result;
  };
  const auto max3 = [](double a, double b, double c)
  {
    double result = a;
    if (result < b) { result = b; }
    if (result < c) { result = c; }
    return result;
  };
  const V3 low_grid = Transform(frame.x_gw, V3{min3(v[0].x, v[1].x, v[2].x),
                                               min3(v[0].y, v[1].y, v[2].y),
                                               min3(v[0].z, v[1].z, v[2].z)});
  const V3 high_grid = Transform(frame.x_gw, V3{max3(v[0].x, v[1].x, v[2].x),
                                                max3(v[0].y, v[1].y, v[2].y),
                                                max3(v[0].z, v[1].z, v[2].z)});
  int64_t lo[3] = {static_cast<int64_t>(floor(low_grid.x * frame.inverse_resolution)),
                   static_cast<int64_t>(floor(low_grid.y * frame.inverse_resolution)),
                   static_cast<int64_t>(floor(low_grid.z * frame.inverse_resolution))};
  int64_t hi[3] = {static_cast<int64_t>(floor(high_grid.x * frame.inverse_resolution)),
                   static_cast<int64_t>(floor(high_grid.y * frame.inverse_resolution)),
                   static_cast<int64_t>(floor(high_grid.z * frame.inverse_resolution))};
  if (!enforce_contains)
  {
    // cells outside the map have no effect: walk only the part of the box inside it
    lo[0] = max(lo[0], int64_t{0});
    lo[1] = max(lo[1], int64_t{0});
    lo[2] = max(lo[2], int64_t{0});
    hi[0] = min(hi[0], frame.nx - 1);
    hi[1] = min(hi[1], frame.ny - 1);
    hi[2] = min(hi[2], frame.nz - 1);
  }
  if (hi[0] < lo[0] || hi[1] < lo[1] || hi[2] < lo[2])
  {
    return;
  }
  const int64_t span_y = hi[1] - lo[1] + 1;
  const int64_t span_z = hi[2] - lo[2] + 1;
  const int64_t box_cells = (hi[0] - lo[0] + 1) * span_y * span_z;
  for (int64_t i = static_cast<int64_t>(blockIdx.y) * kThreads + threadIdx.x; i < box_cells;
       i += static_cast<int64_t>(kThreads) * gridDim.y)
  {
    const int64_t z = lo[2] + i % span_z;
    const int64_t y = lo[1] + (i / span_z) % span_y;
    const int64_t x = lo[0] + i / (span_z * span_y);
    const V3 centre{frame.resolution * (static_cast<double>(x) + 0.5),
                    frame.resolution * (static_cast<double>(y) + 0.5),
                    frame.resolution * (static_cast<double>(z) + 0.5)};
    const V3 q = Transform(frame.x_wg, centre);
    const V3 offset = Sub(ClosestOnTriangle(v[0], v[1], v[2], normal, q), q);
    if (Dot(offset, offset) <= frame.max_check_radius_squared)
    {
      if (x >= 0 && x < frame.nx && y >= 0 && y < frame.ny && z >= 0 && z < frame.nz)
      {
        cells[((x * frame.ny + y) * frame.nz + z) * cell_words] = 1.0f;
      }
      else if (enforce_contains)
      {
        atomicOr(flags, kFlagNotContained);
      }
    }
  }
}

int MakeFrame(int64_t nx, int64_t ny, int64_t nz, double resolution, const double* x_wg,
              const double* x_gw, Frame* frame)
{
  if (!ValidDims(nx, ny, nz))
  {
    return FailInvalid("grid dimensions %lld x %lld x %lld out of range",
                       static_cast<long long>(nx), static_cast<long long>(ny),
                       static_cast<long long>(nz));
  }
  if (!(resolution > 0.0) || !std::isfinite(resolution))
  {
    return FailInvalid("resolution must be greater than zero");
  }
  if (x_wg == nullptr || x_gw == nullptr)
  {
    return FailInvalid("null origin transform");
  }
  for (int i = 0; i < 16; i++)
  {
    frame->x_wg[i] = x_wg[i];
    frame->x_gw[i] = x_gw[i];
  }
  frame->nx = nx;
  frame->ny = ny;
  frame->nz = nz;
  frame->resolution = resolution;
  frame->inverse_resolution = 1.0 / resolution;
  // mesh_rasterizer.cpp:118-120
  const double min_check_radius = resolution * 0.5;
  const double max_check_radius = min_check_radius * std::sqrt(3.0);
  frame->max_check_radius_squared = std::pow(max_check_radius, 2.0);
  return VGT_B200_OK;
}

int CheckMesh(const void* vertices, int64_t num_vertices, const void* triangles,
              int64_t num_triangles, const void* cells, int cell_bytes)
{
  if (num_vertices < 0 || num_triangles < 0 || num_triangles > 0x7fffffffLL
      || (num_vertices > 0 && vertices == nullptr) || (num_triangles > 0 && triangles == nullptr))
  {
    return FailInvalid("invalid mesh arrays");
  }
  if (cells == nullptr)
  {
    return FailInvalid("null cell array");
  }
  if (cell_bytes != 4 && cell_bytes != 8)
  {
    return FailInvalid("cell size must be 4 (OccupancyCell) or 8 (OccupancyComponentCell) bytes");
  }
  return VGT_B200_OK;
}

int FlagsToStatus(int flags)
{
  if (flags & kFlagBadIndex)
  {
    SetLastError("a triangle names a vertex that does not exist");
    return VGT_B200_ERR_OUT_OF_RANGE;
  }
  if (flags & kFlagNotContained)
  {
    SetLastError("Triangle is not contained by occupancy map");
    return VGT_B200_ERR_NOT_CONTAINED;
  }
  return VGT_B200_OK;
}

int LaunchRasterize(const double* d_vertices, int64_t num_vertices, const int32_t* d_triangles,
                    int64_t num_triangles, const Frame& frame, int enforce_contains,
                    float* d_cells, int cell_words, int* d_flags, cudaStream_t stream)
{
  VGT_CUDA_TRY(cudaMemsetAsync(d_flags, 0, sizeof(int), stream), "flag reset");
  if (num_triangles > 0)
  {
    const dim3 grid(static_cast<unsigned>(num_triangles), kSlices);
    RasterizeTrianglesKernel<<<grid, kThreads, 0, stream>>>(
        d_vertices, num_vertices, d_triangles, frame, enforce_contains, d_cells, cell_words,
        d_flags); NoteKernelLaunch();
    VGT_CUDA_TRY(cudaGetLastError(), "RasterizeTrianglesKernel launch");
  }
  return VGT_B200_OK;
}
}  // namespace
}  // namespace rasterizer
}  // namespace vgt_b200

using namespace vgt_b200;
using namespace vgt_b200::rasterizer;

extern "C"
{
int vgt_b200_rasterize_mesh_dev(const double* d_vertices_xyz, int64_t num_vertices,
                                const int32_t* d_triangles, int64_t num_triangles, void* d_cells,
                                int cell_bytes, int64_t nx, int64_t ny, int64_t nz,
                                double resolution, const double* x_wg, const double* x_gw,
                                int enforce_contains, int device, int* d_flags, void* stream)
{
  const int check = CheckMesh(d_vertices_xyz, num_vertices, d_triangles, num_triangles, d_cells,
                              cell_bytes);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  if (d_flags == nullptr)
  {
    return FailInvalid("null flag word");
  }
  Frame frame;
  const int framed = MakeFrame(nx, ny, nz, resolution, x_wg, x_gw, &frame);
  if (framed != VGT_B200_OK)
  {
    return framed;
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  return LaunchRasterize(d_vertices_xyz, num_vertices, d_triangles, num_triangles, frame,
                         enforce_contains, static_cast<float*>(d_cells), cell_bytes / 4, d_flags,
                         static_cast<cudaStream_t>(stream));
}

int vgt_b200_rasterize_status(int flags)
{
  return FlagsToStatus(flags);
}

int vgt_b200_rasterize_mesh_f64(const double* vertices_xyz, int64_t num_vertices,
                                const int32_t* triangles, int64_t num_triangles, void* cells,
                                int cell_bytes, int64_t nx, int64_t ny, int64_t nz,
                                double resolution, const double* x_wg, const double* x_gw,
                                int enforce_contains, int device)
{
  const int check = CheckMesh(vertices_xyz, num_vertices, triangles, num_triangles, cells,
                              cell_bytes);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  Frame frame;
  const int framed = MakeFrame(nx, ny, nz, resolution, x_wg, x_gw, &frame);
  if (framed != VGT_B200_OK)
  {
    return framed;
  }
  if (num_triangles == 0)
  {
    return VGT_B200_OK;
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  KeepPoolMemory(device);
  cudaStream_t stream = nullptr;
  VGT_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
  struct StreamOwner
  {
    cudaStream_t stream;
    ~StreamOwner()
    {
      cudaStreamSynchronize(stream);
      cudaStreamDestroy(stream);
    }
  } owner{stream};
  const size_t cell_count = static_cast<size_t>(nx * ny * nz);
  const size_t grid_bytes = cell_count * static_cast<size_t>(cell_bytes);
  int flags = 0;
  {
    // (declared inside the stream owner's scope: freed on the stream before it is destroyed)
    StreamScratch<double> d_vertices;
    StreamScratch<int32_t> d_triangles;
    StreamScratch<char> d_cells;
    StreamScratch<int> d_flags;
    VGT_CUDA_TRY(d_vertices.Allocate(std::max<int64_t>(1, num_vertices * 3), stream),
                 "vertex allocation");
    VGT_CUDA_TRY(d_triangles.Allocate(num_triangles * 3, stream), "triangle allocation");
    VGT_CUDA_TRY(d_cells.Allocate(static_cast<int64_t>(grid_bytes), stream), "cell allocation");
    VGT_CUDA_TRY(d_flags.Allocate(1, stream), "flag allocation");
    VGT_CUDA_TRY(cudaMemcpyAsync(d_vertices.get(), vertices_xyz,
                                 sizeof(double) * 3 * static_cast<size_t>(num_vertices),
                                 cudaMemcpyHostToDevice, stream),
                 "copy vertices to device");
    VGT_CUDA_TRY(cudaMemcpyAsync(d_triangles.get(), triangles,
                                 sizeof(int32_t) * 3 * static_cast<size_t>(num_triangles),
                                 cudaMemcpyHostToDevice, stream),
                 "copy triangles to device");
    StagedTransfer transfer;
    VGT_CUDA_TRY(transfer.ToDevice(d_cells.get(), grid_bytes, static_cast<const char*>(cells),
                                   grid_bytes, grid_bytes, 1, stream),
                 "copy cells to device");
    const int launched = LaunchRasterize(
        d_vertices.get(), num_vertices, d_triangles.get(), num_triangles, frame, enforce_contains,
        reinterpret_cast<float*>(d_cells.get()), cell_bytes / 4, d_flags.get(), stream);
    if (launched != VGT_B200_OK)
    {
      cudaStreamSynchronize(stream);
      return launched;
    }
    VGT_CUDA_TRY(cudaMemcpyAsync(&flags, d_flags.get(), sizeof(int), cudaMemcpyDeviceToHost,
                                 stream),
                 "copy flags to host");
    VGT_CUDA_TRY(transfer.ToHost(static_cast<char*>(cells), grid_bytes, d_cells.get(), grid_bytes,
                                 grid_bytes, 1, stream),
                 "copy cells to host");
    VGT_CUDA_TRY(cudaStreamSynchronize(stream), "mesh rasterization");
  }
  return FlagsToStatus(flags);
}
}  // extern "C"
