// =============================================================================
// TEST INFRASTRUCTURE ONLY -- CPU oracle for the raycasting voxelizer.
//
// From-scratch restatement of the reference CPU voxelizer. Only tests/,
// bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke()
// may use it; the product path never does.
//
// Parity status: the reference's own tests pin only coarse outcomes for this
// part (plane-level occupancy on an 8^3 scene, test/pointcloud_voxelization_test.cpp:84-158,
// and the "each voxel at most once per ray" property, test/voxel_raycasting_test.cpp:61-100);
// this oracle is checked against both. Raw seen-free / seen-filled counts and
// the last bits of the double DDA are NOT pinned by any reference fixture:
// count parity is defined against this file's fixed operation order
// (left-to-right sums, no FMA contraction: build with -ffp-contract=off).
//
// What it follows (paths relative to the reference checkout):
//   src/voxelized_geometry_tools/cpu_pointcloud_voxelization.cpp:133-165  driver
//   src/voxelized_geometry_tools/cpu_pointcloud_voxelization.cpp:167-206  per cloud
//   src/voxelized_geometry_tools/cpu_pointcloud_voxelization.cpp:208-436  per ray DDA
//   src/voxelized_geometry_tools/cpu_pointcloud_voxelization.cpp:438-497  filter
//   include/voxelized_geometry_tools/pointcloud_voxelization_interface.hpp:55-86
//       per-camera seen-as rule
//   include/voxelized_geometry_tools/cpu_pointcloud_voxelization.hpp:24-32
//       tracking cell = {seen_free, seen_filled} int32 pair
//
// Third-party arithmetic restated here (common_robotics_utilities VoxelGrid and
// Eigen, neither vendored in the reference; conventions confirmed by the
// in-tree float kernel, src/voxelized_geometry_tools/cuda_voxelization_helpers.cu):
//   location -> index : floor(p * (1.0 / voxel_size))            (cuda.cu:139-144)
//   index -> centre   : voxel_size * (index + 0.5)               (cuda.cu:250-255)
//   grid extent       : voxel_count * voxel_size
//   X_GC * p          : ((m0*x + m1*y) + m2*z) + m3, per row
//   |ray|             : sqrt((x*x + y*y) + z*z)
// =============================================================================
#include <atomic>
#include <cmath>
#include <cstdint>
#include <limits>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace
{
constexpr double kPosInf = std::numeric_limits<double>::infinity();

struct Vec3
{
  double x;
  double y;
  double z;
  double operator[](int axis) const
  {
    return (axis == 0) ? x : ((axis == 1) ? y : z);
  }
};

struct Cell3
{
  int64_t x;
  int64_t y;
  int64_t z;
  bool operator==(const Cell3& o) const
  {
    return x == o.x && y == o.y && z == o.z;
  }
  bool operator!=(const Cell3& o) const { return !(*this == o); }
};

struct TrackingGrid
{
  int64_t nx;
  int64_t ny;
  int64_t nz;
  double voxel_size;
  double inverse_voxel_size;
  int32_t* counts;  // [voxel][0] = seen free, [voxel][1] = seen filled

  bool Contains(const Cell3& c) const
  {
    return c.x >= 0 && c.x < nx && c.y >= 0 && c.y < ny && c.z >= 0
        && c.z < nz;
  }
  Cell3 CellOf(const Vec3& p) const
  {
    return Cell3{
        static_cast<int64_t>(std::floor(p.x * inverse_voxel_size)),
        static_cast<int64_t>(std::floor(p.y * inverse_voxel_size)),
        static_cast<int64_t>(std::floor(p.z * inverse_voxel_size))};
  }
  Vec3 CentreOf(const Cell3& c) const
  {
    return Vec3{voxel_size * (static_cast<double>(c.x) + 0.5),
                voxel_size * (static_cast<double>(c.y) + 0.5),
                voxel_size * (static_cast<double>(c.z) + 0.5)};
  }
  Vec3 Extent() const
  {
    return Vec3{static_cast<double>(nx) * voxel_size,
                static_cast<double>(ny) * voxel_size,
                static_cast<double>(nz) * voxel_size};
  }
  void Bump(const Cell3& c, int which) const
  {
    int32_t* slot = counts + 2 * ((c.x * ny + c.y) * nz + c.z) + which;
#ifdef _OPENMP
#pragma omp atomic
#endif
    *slot += 1;
  }
};

inline int StepToward(int64_t difference)
{
  return (difference > 0) ? 1 : ((difference < 0) ? -1 : 0);
}

// cpu_pcv.cpp:336-354.
inline double FirstBoundaryT(
    double point_axis, double ray_axis, double cell_low, double cell_high)
{
  if (ray_axis > 0.0)
  {
    return std::abs((cell_high - point_axis) / ray_axis);
  }
  else if (ray_axis < -0.0)
  {
    return std::abs((point_axis - cell_low) / ray_axis);
  }
  return kPosInf;
}

// cpu_pcv.cpp:208-436, one ray from the sensor origin to one point.
void CastOneRay(
    const Vec3& origin, const Cell3& origin_cell, const Vec3& point,
    double max_range, const TrackingGrid& grid)
{
  // Step 1: clip to max range (:217-226).
  const Vec3 ray{point.x - origin.x, point.y - origin.y, point.z - origin.z};
  const double ray_length =
      std::sqrt((ray.x * ray.x + ray.y * ray.y) + ray.z * ray.z);
  const bool clipped = ray_length > max_range;
  Vec3 final_point = point;
  if (clipped)
  {
    const double scale = max_range / ray_length;
    final_point = Vec3{origin.x + ray.x * scale, origin.y + ray.y * scale,
                       origin.z + ray.z * scale};
  }

  // Step 2: if the origin cell is outside, enter through the slabs (:229-290).
  Vec3 start_point = origin;
  if (!grid.Contains(origin_cell))
  {
    const Vec3 extent = grid.Extent();
    double t_enter = 0.0;
    double t_exit = max_range;
    const Vec3 direction{
        ray.x / ray_length, ray.y / ray_length, ray.z / ray_length};
    const double flat_threshold = 1e-10;
    for (int axis = 0; axis < 3; axis++)
    {
      if (std::abs(direction[axis]) < flat_threshold)
      {
        const bool inside_slab =
            origin[axis] >= 0.0 && origin[axis] < extent[axis];
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
        // Literal mirror of cpu_pcv.cpp:274-277: the exit bound GROWS.
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
    const double nudge = 1e-10;
    const double advance = t_enter + nudge;
    start_point = Vec3{origin.x + direction.x * advance,
                       origin.y + direction.y * advance,
                       origin.z + direction.z * advance};
  }

  // Steps 3-4: end cells and step signs (:293-321).
  const Cell3 start_cell = grid.CellOf(start_point);
  const Cell3 final_cell = grid.CellOf(final_point);
  const int64_t step_x = StepToward(final_cell.x - start_cell.x);
  const int64_t step_y = StepToward(final_cell.y - start_cell.y);
  const int64_t step_z = StepToward(final_cell.z - start_cell.z);

  // Step 5: first boundary crossing and per-cell increments (:324-365).
  const double half = grid.voxel_size * 0.5;
  const Vec3 centre = grid.CentreOf(start_cell);
  double tx = FirstBoundaryT(start_point.x, ray.x, centre.x - half, centre.x + half);
  double ty = FirstBoundaryT(start_point.y, ray.y, centre.y - half, centre.y + half);
  double tz = FirstBoundaryT(start_point.z, ray.z, centre.z - half, centre.z + half);
  const double dtx = std::abs(grid.voxel_size / ray.x);
  const double dty = std::abs(grid.voxel_size / ray.y);
  const double dtz = std::abs(grid.voxel_size / ray.z);

  // Step 6: the final cell first (:368-381).
  if (grid.Contains(final_cell))
  {
    grid.Bump(final_cell, clipped ? 0 : 1);
  }

  // Walk (:384-435): x wins ties, then y, then z.
  Cell3 at = start_cell;
  while (at != final_cell)
  {
    if (!grid.Contains(at))
    {
      break;
    }
    grid.Bump(at, 0);
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

inline bool AllFinite(double x, double y, double z)
{
  return std::isfinite(x) && std::isfinite(y) && std::isfinite(z);
}

int ResolveThreads(int requested)
{
#ifdef _OPENMP
  return (requested <= 0) ? omp_get_max_threads() : requested;
#else
  (void)requested;
  return 1;
#endif
}
}  // namespace

extern "C"
{
// Raycasts one cloud into `counts` (int32 [nx*ny*nz][2], accumulated, the
// caller zeroes it). `points_xyz` are in the cloud frame; `x_gc` is the 4x4
// grid-from-cloud transform in column-major order (Eigen's .data()).
// Mirrors cpu_pcv.cpp:167-206.
int vgt_oracle_raycast_cloud_f64(
    const double* points_xyz, int64_t num_points, const double* x_gc,
    double max_range, int64_t nx, int64_t ny, int64_t nz, double voxel_size,
    int threads, int32_t* counts)
{
  if (nx < 1 || ny < 1 || nz < 1 || counts == nullptr || x_gc == nullptr)
  {
    return 1;
  }
  const TrackingGrid grid{nx, ny, nz, voxel_size, 1.0 / voxel_size, counts};
  const Vec3 origin{x_gc[12], x_gc[13], x_gc[14]};
  const Cell3 origin_cell = grid.CellOf(origin);
  const int resolved = ResolveThreads(threads);
  (void)resolved;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(resolved)
#endif
  for (int64_t i = 0; i < num_points; i++)
  {
    const double px = points_xyz[3 * i + 0];
    const double py = points_xyz[3 * i + 1];
    const double pz = points_xyz[3 * i + 2];
    if (!AllFinite(px, py, pz))
    {
      continue;
    }
    const Vec3 in_grid{
        ((x_gc[0] * px + x_gc[4] * py) + x_gc[8] * pz) + x_gc[12],
        ((x_gc[1] * px + x_gc[5] * py) + x_gc[9] * pz) + x_gc[13],
        ((x_gc[2] * px + x_gc[6] * py) + x_gc[10] * pz) + x_gc[14]};
    CastOneRay(origin, origin_cell, in_grid, max_range, grid);
  }
  return 0;
}

// One ray given directly in the grid frame (RaycastSinglePoint,
// cpu_pcv.cpp:81-109). Returns 2 for non-finite input like the reference's
// invalid_argument.
int vgt_oracle_raycast_single_f64(
    const double* origin_xyz, const double* point_xyz, double max_range,
    int64_t nx, int64_t ny, int64_t nz, double voxel_size, int32_t* counts)
{
  if (!AllFinite(origin_xyz[0], origin_xyz[1], origin_xyz[2])
      || !AllFinite(point_xyz[0], point_xyz[1], point_xyz[2]))
  {
    return 2;
  }
  const TrackingGrid grid{nx, ny, nz, voxel_size, 1.0 / voxel_size, counts};
  const Vec3 origin{origin_xyz[0], origin_xyz[1], origin_xyz[2]};
  const Vec3 point{point_xyz[0], point_xyz[1], point_xyz[2]};
  CastOneRay(origin, grid.CellOf(origin), point, max_range, grid);
  return 0;
}

// Per-voxel combine + filter over `num_grids` tracking grids laid out as
// counts[grid][voxel][2]; `occupancy` is updated in place.
// Mirrors cpu_pcv.cpp:438-497 and pcv_if.hpp:55-86.
int vgt_oracle_filter_f32(
    const int32_t* counts, int32_t num_grids, int64_t num_voxels,
    double percent_seen_free, int32_t outlier_points_threshold,
    int32_t num_cameras_seen_free, int threads, float* occupancy)
{
  if (percent_seen_free <= 0.0 || percent_seen_free > 1.0
      || outlier_points_threshold <= 0 || num_cameras_seen_free <= 0)
  {
    return 2;
  }
  const int resolved = ResolveThreads(threads);
  (void)resolved;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(resolved)
#endif
  for (int64_t voxel = 0; voxel < num_voxels; voxel++)
  {
    if (!(occupancy[voxel] <= 0.5))
    {
      continue;
    }
    int32_t cameras_free = 0;
    int32_t cameras_filled = 0;
    for (int32_t g = 0; g < num_grids; g++)
    {
      const int32_t* cell = counts + 2 * (g * num_voxels + voxel);
      const int32_t seen_free = cell[0];
      const int32_t seen_filled =
          (cell[1] >= outlier_points_threshold) ? cell[1] : 0;
      if (seen_free > 0 && seen_filled > 0)
      {
        const double fraction_free = static_cast<double>(seen_free)
            / static_cast<double>(seen_free + seen_filled);
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
    if (cameras_filled > 0)
    {
      occupancy[voxel] = 1.0f;
    }
    else if (cameras_free >= num_cameras_seen_free)
    {
      occupancy[voxel] = 0.0f;
    }
    else
    {
      occupancy[voxel] = 0.5f;
    }
  }
  return 0;
}
}  // extern "C"
