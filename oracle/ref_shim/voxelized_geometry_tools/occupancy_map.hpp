// TEST INFRASTRUCTURE ONLY -- stand-in that SHADOWS the reference's occupancy_map.hpp (which needs
// common_robotics_utilities serialization / maybe / utility) with the cell layout and grid surface
// the voxelizer interface and our adapter use (reference lines 28-58, 65-67, 160-216).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include <Eigen/Geometry>
#include <common_robotics_utilities/utility.hpp>
#include <common_robotics_utilities/voxel_grid.hpp>
#include <voxelized_geometry_tools/signed_distance_field.hpp>
#include <voxelized_geometry_tools/vgt_namespace.hpp>

namespace voxelized_geometry_tools
{
VGT_NAMESPACE_BEGIN
class OccupancyCell
{
public:
  OccupancyCell() = default;
  explicit OccupancyCell(const float occupancy) : occupancy_(occupancy) {}
  float Occupancy() const { return occupancy_; }
  void SetOccupancy(const float occupancy) { occupancy_ = occupancy; }

private:
  float occupancy_ = 0.0f;
};
static_assert(sizeof(OccupancyCell) == sizeof(float), "OccupancyCell is larger than expected.");

class OccupancyMap
    : public common_robotics_utilities::voxel_grid::VoxelGridBase<
          OccupancyCell, std::vector<OccupancyCell>>
{
public:
  using Base = common_robotics_utilities::voxel_grid::VoxelGridBase<
      OccupancyCell, std::vector<OccupancyCell>>;
  OccupancyMap() = default;
  OccupancyMap(const Eigen::Isometry3d& origin_transform, const std::string& frame,
               const common_robotics_utilities::voxel_grid::VoxelGridSizes& sizes,
               const OccupancyCell& default_value)
      : Base(origin_transform, sizes, default_value), frame_(frame) {}
  double Resolution() const { return VoxelXSize(); }
  const std::string& Frame() const { return frame_; }

private:
  std::string frame_;
};
VGT_NAMESPACE_END
}  // namespace voxelized_geometry_tools
