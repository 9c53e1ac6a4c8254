// TEST INFRASTRUCTURE ONLY -- shared body of the stand-ins for the reference's two component-map
// headers (occupancy_component_map.hpp, tagged_object_occupancy_component_map.hpp). The real
// headers need the logging / topology layers of common_robotics_utilities; the C++ adapter only
// needs the packed cell layouts (which the reference pins with static_asserts) and the grid
// surface. (OccupancyMap and TaggedObjectOccupancyMap are the reference's own classes: their
// headers compile over the shim as they are.)
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
namespace stand_in
{
template <typename CellType>
class CellMap
    : public common_robotics_utilities::voxel_grid::VoxelGridBase<CellType, std::vector<CellType>>
{
public:
  using Base =
      common_robotics_utilities::voxel_grid::VoxelGridBase<CellType, std::vector<CellType>>;
  CellMap() = default;
  CellMap(const Eigen::Isometry3d& origin_transform, const std::string& frame,
          const common_robotics_utilities::voxel_grid::VoxelGridSizes& sizes,
          const CellType& default_value)
      : Base(origin_transform, sizes, default_value), frame_(frame) {}
  double Resolution() const { return this->VoxelXSize(); }
  const std::string& Frame() const { return frame_; }

private:
  std::string frame_;
};
}  // namespace stand_in
VGT_NAMESPACE_END
}  // namespace voxelized_geometry_tools
