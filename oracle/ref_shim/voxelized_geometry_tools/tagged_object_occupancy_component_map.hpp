// TEST INFRASTRUCTURE ONLY -- stand-in that SHADOWS the reference's
// tagged_object_occupancy_component_map.hpp: the cell layout of its lines 29-100 (float
// occupancy, uint32 object id, uint32 component, uint32 spatial segment, 16 bytes) and the grid
// surface the C++ adapter uses.
#pragma once

#include <voxelized_geometry_tools/cell_map_stand_in.hpp>

namespace voxelized_geometry_tools
{
VGT_NAMESPACE_BEGIN
class TaggedObjectOccupancyComponentCell
{
public:
  TaggedObjectOccupancyComponentCell() = default;
  explicit TaggedObjectOccupancyComponentCell(const float occupancy) : occupancy_(occupancy) {}
  TaggedObjectOccupancyComponentCell(const float occupancy, const uint32_t object_id)
      : occupancy_(occupancy), object_id_(object_id) {}
  TaggedObjectOccupancyComponentCell(const float occupancy, const uint32_t object_id,
                                     const uint32_t component, const uint32_t spatial_segment)
      : occupancy_(occupancy), object_id_(object_id), component_(component),
        spatial_segment_(spatial_segment) {}
  float Occupancy() const { return occupancy_.load(); }
  uint32_t ObjectId() const { return object_id_.load(); }
  uint32_t Component() const { return component_.load(); }
  uint32_t SpatialSegment() const { return spatial_segment_.load(); }
  void SetOccupancy(const float occupancy) { occupancy_.store(occupancy); }
  void SetObjectId(const uint32_t object_id) { object_id_.store(object_id); }

private:
  common_robotics_utilities::utility::CopyableMoveableAtomic<float, std::memory_order_relaxed>
      occupancy_{0.0f};
  common_robotics_utilities::utility::CopyableMoveableAtomic<uint32_t, std::memory_order_relaxed>
      object_id_{0u};
  common_robotics_utilities::utility::CopyableMoveableAtomic<uint32_t, std::memory_order_relaxed>
      component_{0u};
  common_robotics_utilities::utility::CopyableMoveableAtomic<uint32_t, std::memory_order_relaxed>
      spatial_segment_{0u};
};
static_assert(sizeof(TaggedObjectOccupancyComponentCell) == (sizeof(float) * 4),
              "TaggedObjectOccupancyComponentCell is larger than expected.");

using TaggedObjectOccupancyComponentMap = stand_in::CellMap<TaggedObjectOccupancyComponentCell>;
VGT_NAMESPACE_END
}  // namespace voxelized_geometry_tools
