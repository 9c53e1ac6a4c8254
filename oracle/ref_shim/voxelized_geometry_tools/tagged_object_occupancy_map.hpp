// TEST INFRASTRUCTURE ONLY -- stand-in that SHADOWS the reference's
// tagged_object_occupancy_map.hpp: the cell layout of its lines 29-72 (two relaxed atomics,
// float occupancy then uint32 object id, 8 bytes) and the grid surface the C++ adapter uses.
#pragma once

#include <voxelized_geometry_tools/cell_map_stand_in.hpp>

namespace voxelized_geometry_tools
{
VGT_NAMESPACE_BEGIN
class TaggedObjectOccupancyCell
{
public:
  TaggedObjectOccupancyCell() : occupancy_(0.0f), object_id_(0u) {}
  explicit TaggedObjectOccupancyCell(const float occupancy)
      : occupancy_(occupancy), object_id_(0u) {}
  TaggedObjectOccupancyCell(const float occupancy, const uint32_t object_id)
      : occupancy_(occupancy), object_id_(object_id) {}
  float Occupancy() const { return occupancy_.load(); }
  uint32_t ObjectId() const { return object_id_.load(); }
  void SetOccupancy(const float occupancy) { occupancy_.store(occupancy); }
  void SetObjectId(const uint32_t object_id) { object_id_.store(object_id); }

private:
  common_robotics_utilities::utility::CopyableMoveableAtomic<float, std::memory_order_relaxed>
      occupancy_{0.0f};
  common_robotics_utilities::utility::CopyableMoveableAtomic<uint32_t, std::memory_order_relaxed>
      object_id_{0u};
};
static_assert(sizeof(TaggedObjectOccupancyCell) == (sizeof(float) * 2),
              "TaggedObjectOccupancyCell is larger than expected.");

using TaggedObjectOccupancyMap = stand_in::CellMap<TaggedObjectOccupancyCell>;
VGT_NAMESPACE_END
}  // namespace voxelized_geometry_tools
