// Occupancy cells of the other map types -> filled mask, and the merge of the "free" and "named
// objects" fields (sm_100a). Included by edt_kernels.cu.
//
// The reference's other three map types store 8- or 16-byte cells
//   OccupancyComponentCell             { float occupancy; uint32 component; }
//     (include/voxelized_geometry_tools/occupancy_component_map.hpp:28-64)
//   TaggedObjectOccupancyCell          { float occupancy; uint32 object_id; }
//     (include/.../tagged_object_occupancy_map.hpp:28-68)
//   TaggedObjectOccupancyComponentCell { float occupancy; uint32 object_id; uint32 component;
//                                        uint32 spatial_segment; }
//     (include/.../tagged_object_occupancy_component_map.hpp:17-60)
// and their ExtractSignedDistanceField differs from OccupancyMap's only in the filled predicate
// (occupancy_component_map.hpp:270-306, tagged_object_occupancy_map.hpp:199-247,
// tagged_object_occupancy_component_map.hpp:360-410). The predicate is evaluated here, on the
// device, straight from the raw cell array (read once, 1 byte per voxel out); the mask then
// takes the same three passes as every other grid.
#pragma once

#include <cstdint>

#include "edt_device.cuh"

namespace vgt_b200
{
namespace edt
{
namespace
{
// What a cell must satisfy, besides the occupancy rule, to count as filled.
enum ObjectRule
{
  kAnyObject = 0,     // no object restriction (objects_to_use empty / maps without object ids)
  kListedObjects = 1, // object_id in the sorted list (tagged_object_occupancy_map.hpp:219-222)
  kNamedObjects = 2   // object_id > 0 (tagged_object_occupancy_map.hpp:321-324)
};

// cells: raw array of `cell_words` 32-bit words per voxel; word 0 = occupancy (float), word 1 =
// object id (only read for the object rules). sorted_ids: ascending, num_ids entries.
__global__ void CellsToMaskKernel(const uint32_t* __restrict__ cells, int cell_words,
                                  int64_t count, int unknown_is_filled, int object_rule,
                                  const uint32_t* __restrict__ sorted_ids, int num_ids,
                                  uint8_t* __restrict__ mask)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count)
  {
    return;
  }
  const uint32_t* cell = cells + i * cell_words;
  bool filled = IsFilled(__uint_as_float(__ldcs(cell)), unknown_is_filled);
  if (filled && object_rule != kAnyObject)
  {
    const uint32_t object_id = __ldcs(cell + 1);
    if (object_rule == kNamedObjects)
    {
      filled = object_id > 0u;
    }
    else
    {
      int low = 0;
      int high = num_ids;
      while (low < high)
      {
        const int middle = (low + high) >> 1;
        if (__ldg(sorted_ids + middle) < object_id)
        {
          low = middle + 1;
        }
        else
        {
          high = middle;
        }
      }
      filled = (low < num_ids) && (__ldg(sorted_ids + low) == object_id);
    }
  }
  mask[i] = filled ? 1 : 0;
}

// ExtractFreeAndNamedObjectsSignedDistanceField's merge (tagged_object_occupancy_map.hpp:344-369):
//   free >= 0 -> free;  else named <= -0 -> named;  else 0.   Then Lock()'s min/max.
// Grid-stride: a fixed grid, so the min/max atomics are two per warp of the grid, not two per
// 32 voxels (all to the same two addresses, which the L2 serialises).
template <typename Out, typename Key>
__global__ void MergeFreeAndNamedKernel(const Out* __restrict__ free_sdf,
                                        const Out* __restrict__ named_sdf, int64_t count,
                                        Out* __restrict__ combined, Key* min_max_keys)
{
  Out low = PositiveInfinity<Out>();
  Out high = -PositiveInfinity<Out>();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += stride)
  {
    const Out free_value = free_sdf[i];
    const Out named_value = named_sdf[i];
    const Out value =
        (free_value >= Out(0)) ? free_value : ((named_value <= Out(0)) ? named_value : Out(0));
    combined[i] = value;
    low = (value < low) ? value : low;
    high = (value > high) ? value : high;
  }
  Key key_min = OrderedKey(low);
  Key key_max = OrderedKey(high);
#pragma unroll
  for (int offset = 16; offset > 0; offset >>= 1)
  {
    const Key other_min = __shfl_xor_sync(0xffffffffu, key_min, offset);
    const Key other_max = __shfl_xor_sync(0xffffffffu, key_max, offset);
    key_min = (other_min < key_min) ? other_min : key_min;
    key_max = (other_max > key_max) ? other_max : key_max;
  }
  if ((threadIdx.x & 31) == 0 && min_max_keys != nullptr)
  {
    atomicMin(min_max_keys + 0, key_min);
    atomicMax(min_max_keys + 1, key_max);
  }
}
}  // namespace
}  // namespace edt
}  // namespace vgt_b200
