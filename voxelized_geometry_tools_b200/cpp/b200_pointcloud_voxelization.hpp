// C++ host adapter: a PointCloudVoxelizationInterface backend over the B200 C-ABI.
//
// Takes the place of CudaPointCloudVoxelizer
// (include/voxelized_geometry_tools/device_pointcloud_voxelization.hpp:67-73): same constructor
// shape (options map + logging function), same override
// (pointcloud_voxelization_interface.hpp:295-299). Unlike DevicePointCloudVoxelizer it does not go
// through the float-only DeviceVoxelizationHelperInterface (device_voxelization_interface.hpp:
// 149-156): points are pulled as doubles (CopyPointLocationIntoDoublePtr, pcv_if.hpp:161-166) and
// X_GC is passed as 16 doubles, so its counts equal the CPU backend's.
#pragma once

#include <cstdint>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include <Eigen/Geometry>
#include <voxelized_geometry_tools/occupancy_map.hpp>
#include <voxelized_geometry_tools/pointcloud_voxelization_interface.hpp>
#include <voxelized_geometry_tools/vgt_namespace.hpp>

#include "vgt_b200.h"

namespace voxelized_geometry_tools
{
VGT_NAMESPACE_BEGIN
namespace pointcloud_voxelization
{
class B200PointCloudVoxelizer : public PointCloudVoxelizationInterface
{
public:
  using LoggingFunction = std::function<void(const std::string&)>;

  explicit B200PointCloudVoxelizer(
      const std::map<std::string, int32_t>& options, const LoggingFunction& logging_fn = {})
  {
    // Same option name the reference's CUDA helper reads (cuda_voxelization_helpers.cu:566-590).
    const auto found = options.find("CUDA_DEVICE");
    device_ = (found != options.end()) ? found->second : 0;
    if (logging_fn)
    {
      logging_fn("Option [CUDA_DEVICE] = " + std::to_string(device_));
    }
    // EnforceAvailable (device_pointcloud_voxelization.hpp:34-46): unusable -> runtime_error.
    const int devices = vgt_b200_device_count();
    if (device_ < 0 || device_ >= devices)
    {
      throw std::runtime_error(
          "B200PointCloudVoxelizer: CUDA device " + std::to_string(device_) + " is not available ("
          + std::to_string(devices) + " usable device(s)); there is no CPU fallback");
    }
  }

private:
  VoxelizerRuntime DoVoxelizePointClouds(
      const OccupancyMap& static_environment,
      const PointCloudVoxelizationFilterOptions& filter_options,
      const std::vector<PointCloudWrapperSharedPtr>& pointclouds,
      OccupancyMap& output_environment) const override
  {
    const Eigen::Isometry3d X_GW = static_environment.InverseOriginTransform();
    std::vector<std::vector<double>> points(pointclouds.size());
    std::vector<vgt_b200_cloud> clouds(pointclouds.size());
    for (size_t idx = 0; idx < pointclouds.size(); idx++)
    {
      const PointCloudWrapper& cloud = *pointclouds.at(idx);
      const int64_t size = cloud.Size();
      points[idx].resize(static_cast<size_t>(size) * 3);
      for (int64_t point = 0; point < size; point++)
      {
        cloud.CopyPointLocationIntoDoublePtr(point, points[idx].data() + point * 3);
      }
      // X_GC = X_GW * X_WC (cpu_pointcloud_voxelization.cpp:172-176), column-major 4x4.
      const Eigen::Isometry3d X_GC = X_GW * cloud.PointCloudOriginTransform();
      const double* matrix = X_GC.data();
      for (int i = 0; i < 16; i++)
      {
        clouds[idx].x_gc[i] = matrix[i];
      }
      clouds[idx].points_xyz = points[idx].data();
      clouds[idx].num_points = size;
      clouds[idx].max_range = cloud.MaxRange();
    }
    vgt_b200_filter_options options;
    options.percent_seen_free = filter_options.PercentSeenFree();
    options.outlier_points_threshold = filter_options.OutlierPointsThreshold();
    options.num_cameras_seen_free = filter_options.NumCamerasSeenFree();
    double seconds[2] = {0.0, 0.0};
    // OccupancyCell is one float (occupancy_map.hpp:56-58); device_pointcloud_voxelization.cpp:
    // 165,173 hands the raw data to its helper the same way.
    const float* static_raw =
        reinterpret_cast<const float*>(static_environment.GetImmutableRawData().data());
    float* output_raw = reinterpret_cast<float*>(output_environment.GetMutableRawData().data());
    const int status = vgt_b200_voxelize_f64(
        static_raw, static_environment.NumXVoxels(), static_environment.NumYVoxels(),
        static_environment.NumZVoxels(), static_environment.VoxelXSize(), clouds.data(),
        static_cast<int32_t>(clouds.size()), &options, device_, output_raw, nullptr, seconds);
    if (status == VGT_B200_ERR_INVALID_ARGUMENT)
    {
      throw std::invalid_argument(vgt_b200_last_error());
    }
    if (status != VGT_B200_OK)
    {
      throw std::runtime_error(vgt_b200_last_error());
    }
    return VoxelizerRuntime(seconds[0], seconds[1]);
  }

  int32_t device_ = 0;
};
}  // namespace pointcloud_voxelization
VGT_NAMESPACE_END
}  // namespace voxelized_geometry_tools
