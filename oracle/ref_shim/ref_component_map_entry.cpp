// TEST INFRASTRUCTURE ONLY -- C entry points over the reference's own component map types
// (occupancy_component_map.hpp:270-306 and tagged_object_occupancy_component_map.hpp:360-575 with
// their .cpp files, all unmodified; compiled like ref_map_files_entry.cpp with the reference's
// include directory in front of the shim, into oracle/_ref/libvgt_ref_maps.so):
//   kind 1: OccupancyComponentMap::ExtractSignedDistanceField<T>
//   kind 2: TaggedObjectOccupancyComponentMap::ExtractSignedDistanceField<T>(objects_to_use, ..)
//   kind 3: TaggedObjectOccupancyComponentMap::ExtractFreeAndNamedObjectsSignedDistanceField<T>
// Used by tests/test_oracle_other_maps.py (the C++ adapter test still uses the stand-in headers
// for these two classes).
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include <voxelized_geometry_tools/occupancy_component_map.hpp>
#include <voxelized_geometry_tools/tagged_object_occupancy_component_map.hpp>

namespace
{
namespace vgt = voxelized_geometry_tools;
using common_robotics_utilities::parallelism::DegreeOfParallelism;
using common_robotics_utilities::voxel_grid::Vector3i64;
using common_robotics_utilities::voxel_grid::VoxelGridSizes;

template <typename Scalar>
vgt::SignedDistanceFieldGenerationParameters<Scalar> Parameters(int unknown_is_filled,
                                                                 int add_virtual_border)
{
  return vgt::SignedDistanceFieldGenerationParameters<Scalar>(
      std::numeric_limits<Scalar>::infinity(), DegreeOfParallelism::FromOmp(),
      unknown_is_filled != 0, add_virtual_border != 0);
}

template <typename Scalar>
void CopyOut(const vgt::SignedDistanceField<Scalar>& sdf, void* sdf_out, void* min_max)
{
  std::memcpy(sdf_out, sdf.GetImmutableRawData().data(),
              sizeof(Scalar) * static_cast<size_t>(sdf.NumTotalVoxels()));
  const auto extrema = sdf.GetMinimumMaximum();
  static_cast<Scalar*>(min_max)[0] = extrema.Minimum();
  static_cast<Scalar*>(min_max)[1] = extrema.Maximum();
}

template <typename Scalar>
void Run(int kind, const void* cells, int64_t nx, int64_t ny, int64_t nz, double resolution,
         const std::vector<uint32_t>& objects, int unknown_is_filled, int add_virtual_border,
         void* sdf_out, void* min_max)
{
  const auto sizes = VoxelGridSizes::FromVoxelCounts(resolution, Vector3i64(nx, ny, nz));
  const auto parameters = Parameters<Scalar>(unknown_is_filled, add_virtual_border);
  const size_t count = static_cast<size_t>(nx * ny * nz);
  if (kind == 1)
  {
    static_assert(sizeof(vgt::OccupancyComponentCell) == 8, "packed {float, uint32} cells");
    vgt::OccupancyComponentMap map(Eigen::Isometry3d::Identity(), "reference", sizes,
                                   vgt::OccupancyComponentCell(0.0f, 0u));
    std::memcpy(static_cast<void*>(map.GetMutableRawData().data()), cells,
                sizeof(vgt::OccupancyComponentCell) * count);
    CopyOut(map.template ExtractSignedDistanceField<Scalar>(parameters), sdf_out, min_max);
    return;
  }
  static_assert(sizeof(vgt::TaggedObjectOccupancyComponentCell) == 16, "packed 16-byte cells");
  vgt::TaggedObjectOccupancyComponentMap map(Eigen::Isometry3d::Identity(), "reference", sizes,
                                             vgt::TaggedObjectOccupancyComponentCell());
  std::memcpy(static_cast<void*>(map.GetMutableRawData().data()), cells,
              sizeof(vgt::TaggedObjectOccupancyComponentCell) * count);
  if (kind == 2)
  {
    CopyOut(map.template ExtractSignedDistanceField<Scalar>(objects, parameters), sdf_out,
            min_max);
  }
  else
  {
    CopyOut(map.template ExtractFreeAndNamedObjectsSignedDistanceField<Scalar>(parameters),
            sdf_out, min_max);
  }
}
}  // namespace

extern "C"
{
int vgt_ref_component_map_sdf(int kind, int scalar_bytes, const void* cells, int64_t nx,
                              int64_t ny, int64_t nz, double resolution,
                              const uint32_t* object_ids, int64_t num_object_ids,
                              int unknown_is_filled, int add_virtual_border, void* sdf_out,
                              void* min_max)
{
  try
  {
    if (kind < 1 || kind > 3)
    {
      return 2;
    }
    const std::vector<uint32_t> objects(object_ids, object_ids + num_object_ids);
    if (scalar_bytes == 8)
    {
      Run<double>(kind, cells, nx, ny, nz, resolution, objects, unknown_is_filled,
                  add_virtual_border, sdf_out, min_max);
    }
    else
    {
      Run<float>(kind, cells, nx, ny, nz, resolution, objects, unknown_is_filled,
                 add_virtual_border, sdf_out, min_max);
    }
    return 0;
  }
  catch (...)
  {
    return 1;
  }
}
}  // extern "C"
