// TEST INFRASTRUCTURE ONLY -- stand-in that SHADOWS the reference's occupancy_component_map.hpp:
// the cell layout of its lines 29-70 (float occupancy then uint32 component, 8 bytes) and the
// grid surface the C++ adapter uses.
#pragma once

#include <voxelized_geometry_tools/cell_map_stand_in.hpp>

namespace voxelized_geometry_tools
{
VGT_NAMESPACE_BEGIN
class OccupancyComponentCell
{
public:
  OccupancyComponentCell() : occupancy_(0.0f), component_(0u) {}
  explicit OccupancyComponentCell(const float occupancy) : occupancy_(occupancy), component_(0u) {}
  OccupancyComponentCell(const float occupancy, const uint32_t component)
      : occupancy_(occupancy), component_(component) {}
  float Occupancy() const { return occupancy_.load(); }
  uint32_t Component() const { return component_.load(); }
  void SetOccupancy(const float occupancy) { occupancy_.store(occupancy); }
  void SetComponent(const uint32_t component) { component_.store(component); }

private:
  common_robotics_utilities::utility::CopyableMoveableAtomic<float, std::memory_order_relaxed>
      occupancy_{0.0f};
  common_robotics_utilities::utility::CopyableMoveableAtomic<uint32_t, std::memory_order_relaxed>
      component_{0u};
};
static_assert(sizeof(OccupancyComponentCell) == (sizeof(float) * 2),
              "OccupancyComponentCell is larger than expected.");

using OccupancyComponentMap = stand_in::CellMap<OccupancyComponentCell>;
VGT_NAMESPACE_END
}  // namespace voxelized_geometry_tools
