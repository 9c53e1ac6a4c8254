// TEST INFRASTRUCTURE ONLY -- C entry points over the REFERENCE'S OWN CPU voxelizer.
//
// This translation unit includes the reference's cpu_pointcloud_voxelization.hpp and
// pointcloud_voxelization_interface.hpp unmodified and is linked with the reference's
// cpu_pointcloud_voxelization.cpp compiled unmodified from /root/reference (oracle/Makefile,
// target `ref`). So the ray clipping, the slab test with its tmax quirk, the DDA with its tie
// order and early exits, the final-voxel-first marking and the combine / filter rule that run
// here are literally the reference's (cpu_pointcloud_voxelization.cpp:133-497,
// pointcloud_voxelization_interface.hpp:20-92). Only the third-party layer underneath (Eigen,
// common_robotics_utilities) and the OccupancyMap container are stand-ins (oracle/ref_shim/...);
// the operation order of the small matrix / vector expressions is fixed in
// ref_shim/Eigen/Geometry.
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <voxelized_geometry_tools/cpu_pointcloud_voxelization.hpp>

namespace
{
namespace vgt = voxelized_geometry_tools;
namespace pcv = voxelized_geometry_tools::pointcloud_voxelization;
using common_robotics_utilities::voxel_grid::Vector3i64;
using common_robotics_utilities::voxel_grid::VoxelGridSizes;
using TrackingGrid = pcv::CpuPointCloudVoxelizer::CpuVoxelizationTrackingGrid;
using TrackingCell = pcv::CpuPointCloudVoxelizer::CpuVoxelizationTrackingCell;

// A cloud over caller memory (xyz doubles), the shape of the reference's own test wrapper
// (test/pointcloud_voxelization_test.cpp:23-70).
class ArrayPointCloud : public pcv::PointCloudWrapper
{
public:
  ArrayPointCloud(const double* points, int64_t size, const Eigen::Isometry3d& origin,
                  double max_range)
      : points_(points), size_(size), origin_(origin), max_range_(max_range) {}
  double MaxRange() const override { return max_range_; }
  void SetMaxRange(const double max_range) override { max_range_ = max_range; }
  int64_t Size() const override { return size_; }
  const Eigen::Isometry3d& PointCloudOriginTransform() const override { return origin_; }
  void SetPointCloudOriginTransform(const Eigen::Isometry3d& origin) override { origin_ = origin; }

private:
  void CopyPointLocationIntoDoublePtrImpl(const int64_t index, double* destination) const override
  {
    std::memcpy(destination, points_ + 3 * index, 3 * sizeof(double));
  }
  void CopyPointLocationIntoFloatPtrImpl(const int64_t index, float* destination) const override
  {
    for (int i = 0; i < 3; i++) { destination[i] = static_cast<float>(points_[3 * index + i]); }
  }
  const double* points_;
  int64_t size_;
  Eigen::Isometry3d origin_;
  double max_range_;
};

std::map<std::string, int32_t> Options(int threads)
{
  std::map<std::string, int32_t> options;
  options["CPU_PARALLELIZE"] = 1;
  options["CPU_NUM_THREADS"] = threads > 0 ? threads : -1;
  return options;
}

Eigen::Isometry3d FromColumnMajor(const double* m)
{
  Eigen::Isometry3d out;
  std::memcpy(out.data(), m, 16 * sizeof(double));
  return out;
}

void CopyCounts(const TrackingGrid& grid, int32_t* counts)
{
  const int64_t total = grid.NumTotalVoxels();
  for (int64_t i = 0; i < total; i++)
  {
    const TrackingCell& cell = grid.GetDataIndexImmutable(i);
    counts[2 * i] = cell.seen_free_count.load();
    counts[2 * i + 1] = cell.seen_filled_count.load();
  }
}
}  // namespace

extern "C"
{
// DoVoxelizePointClouds, step by step through the class's public methods so that the raw
// tracking counts can be read back: RaycastPointCloud per cloud (-> DoRaycastPointCloud ->
// DoRaycastSinglePoint), then CombineAndFilterGrids. The grid origin is the identity and cloud c
// is posed at x_gc[c] (column-major 4x4), so the reference's X_GC = X_GW * X_WC is x_gc[c].
//   out_counts: int32 [num_clouds][voxels][2] = {seen_free, seen_filled}, may be null.
int vgt_ref_voxelize_f64(const float* static_occupancy, int64_t nx, int64_t ny, int64_t nz,
                         double voxel_size, int32_t num_clouds, const double* const* points,
                         const int64_t* num_points, const double* x_gc, const double* max_ranges,
                         double percent_seen_free, int32_t outlier_points_threshold,
                         int32_t num_cameras_seen_free, int threads, float* out_occupancy,
                         int32_t* out_counts)
{
  try
  {
    const auto sizes = VoxelGridSizes::FromVoxelCounts(voxel_size, Vector3i64(nx, ny, nz));
    const int64_t total = nx * ny * nz;
    vgt::OccupancyMap output(Eigen::Isometry3d::Identity(), "grid", sizes,
                             vgt::OccupancyCell(0.0f));
    for (int64_t i = 0; i < total; i++)
    {
      output.GetDataIndexMutable(i).SetOccupancy(static_occupancy[i]);
    }
    const pcv::CpuPointCloudVoxelizer voxelizer(Options(threads));
    pcv::CpuPointCloudVoxelizer::VectorCpuVoxelizationTrackingGrid tracking_grids(
        static_cast<size_t>(num_clouds),
        TrackingGrid(Eigen::Isometry3d::Identity(), sizes, TrackingCell()));
    for (int32_t c = 0; c < num_clouds; c++)
    {
      const ArrayPointCloud cloud(points[c], num_points[c], FromColumnMajor(x_gc + 16 * c),
                                  max_ranges[c]);
      voxelizer.RaycastPointCloud(cloud, tracking_grids.at(static_cast<size_t>(c)));
      if (out_counts != nullptr)
      {
        CopyCounts(tracking_grids.at(static_cast<size_t>(c)), out_counts + 2 * total * c);
      }
    }
    const pcv::PointCloudVoxelizationFilterOptions filter_options(
        percent_seen_free, outlier_points_threshold, num_cameras_seen_free);
    voxelizer.CombineAndFilterGrids(filter_options, tracking_grids, output);
    for (int64_t i = 0; i < total; i++)
    {
      out_occupancy[i] = output.GetDataIndexImmutable(i).Occupancy();
    }
    return 0;
  }
  catch (...)
  {
    return 1;
  }
}

// The whole interface call (VoxelizePointClouds -> DoVoxelizePointClouds) with a grid origin
// transform: the reference composes X_GC = X_WG^-1 * X_WC itself.
int vgt_ref_voxelize_posed_f64(const float* static_occupancy, int64_t nx, int64_t ny, int64_t nz,
                               double voxel_size, const double* x_wg, int32_t num_clouds,
                               const double* const* points, const int64_t* num_points,
                               const double* x_wc, const double* max_ranges,
                               double percent_seen_free, int32_t outlier_points_threshold,
                               int32_t num_cameras_seen_free, int threads, float* out_occupancy)
{
  try
  {
    const auto sizes = VoxelGridSizes::FromVoxelCounts(voxel_size, Vector3i64(nx, ny, nz));
    const int64_t total = nx * ny * nz;
    vgt::OccupancyMap static_map(FromColumnMajor(x_wg), "world", sizes, vgt::OccupancyCell(0.0f));
    for (int64_t i = 0; i < total; i++)
    {
      static_map.GetDataIndexMutable(i).SetOccupancy(static_occupancy[i]);
    }
    std::vector<pcv::PointCloudWrapperSharedPtr> clouds;
    for (int32_t c = 0; c < num_clouds; c++)
    {
      clouds.push_back(std::make_shared<ArrayPointCloud>(
          points[c], num_points[c], FromColumnMajor(x_wc + 16 * c), max_ranges[c]));
    }
    const pcv::CpuPointCloudVoxelizer voxelizer(Options(threads));
    const pcv::PointCloudVoxelizationFilterOptions filter_options(
        percent_seen_free, outlier_points_threshold, num_cameras_seen_free);
    const vgt::OccupancyMap result =
        voxelizer.VoxelizePointClouds(static_map, filter_options, clouds);
    for (int64_t i = 0; i < total; i++)
    {
      out_occupancy[i] = result.GetDataIndexImmutable(i).Occupancy();
    }
    return 0;
  }
  catch (...)
  {
    return 1;
  }
}

// RaycastSinglePoint on a caller-provided count grid (accumulates).
int vgt_ref_raycast_single_f64(const double* origin_xyz, const double* point_xyz, double max_range,
                               int64_t nx, int64_t ny, int64_t nz, double voxel_size,
                               int32_t* counts)
{
  try
  {
    const auto sizes = VoxelGridSizes::FromVoxelCounts(voxel_size, Vector3i64(nx, ny, nz));
    TrackingGrid grid(Eigen::Isometry3d::Identity(), sizes, TrackingCell());
    const pcv::CpuPointCloudVoxelizer voxelizer(Options(1));
    voxelizer.RaycastSinglePoint(
        Eigen::Vector4d(origin_xyz[0], origin_xyz[1], origin_xyz[2], 1.0),
        Eigen::Vector4d(point_xyz[0], point_xyz[1], point_xyz[2], 1.0), max_range, grid);
    const int64_t total = nx * ny * nz;
    for (int64_t i = 0; i < total; i++)
    {
      const TrackingCell& cell = grid.GetDataIndexImmutable(i);
      counts[2 * i] += cell.seen_free_count.load();
      counts[2 * i + 1] += cell.seen_filled_count.load();
    }
    return 0;
  }
  catch (...)
  {
    return 1;
  }
}

// out = a * b and out = a^-1 for rigid transforms (column-major 4x4), in the stand-in's
// operation order: what the front ends must reproduce when they compose X_GC.
void vgt_ref_isometry_product(const double* a, const double* b, double* out)
{
  const Eigen::Isometry3d product = FromColumnMajor(a) * FromColumnMajor(b);
  std::memcpy(out, product.data(), 16 * sizeof(double));
}

void vgt_ref_isometry_inverse(const double* a, double* out)
{
  const Eigen::Isometry3d inverse = FromColumnMajor(a).inverse();
  std::memcpy(out, inverse.data(), 16 * sizeof(double));
}
}  // extern "C"
