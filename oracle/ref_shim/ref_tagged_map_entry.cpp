// TEST INFRASTRUCTURE ONLY -- C entry points over the reference's own TaggedObjectOccupancyMap
// (include/voxelized_geometry_tools/tagged_object_occupancy_map.hpp:199-378 and its .cpp, both
// unmodified; compiled like ref_map_files_entry.cpp with the reference's include directory in
// front of the shim, into oracle/_ref/libvgt_ref_maps.so): ExtractSignedDistanceField<T> with
// objects_to_use, and ExtractFreeAndNamedObjectsSignedDistanceField<T>. MakeSeparateObjectSDFs /
// MakeAllObjectSDFs are loops over the former (:249-291).
// Used by tests/test_oracle_other_maps.py and tests/test_gpu_other_maps.py.
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include <voxelized_geometry_tools/tagged_object_occupancy_map.hpp>

namespace
{
namespace vgt = voxelized_geometry_tools;
using common_robotics_utilities::parallelism::DegreeOfParallelism;
using common_robotics_utilities::voxel_grid::Vector3i64;
using common_robotics_utilities::voxel_grid::VoxelGridSizes;

vgt::TaggedObjectOccupancyMap MakeTaggedMap(const void* cells, int64_t nx, int64_t ny, int64_t nz,
                                            double resolution)
{
  static_assert(sizeof(vgt::TaggedObjectOccupancyCell) == 8, "packed {float, uint32} cells");
  const auto sizes = VoxelGridSizes::FromVoxelCounts(resolution, Vector3i64(nx, ny, nz));
  vgt::TaggedObjectOccupancyMap map(Eigen::Isometry3d::Identity(), "reference", sizes,
                                    vgt::TaggedObjectOccupancyCell(0.0f, 0u));
  std::memcpy(static_cast<void*>(map.GetMutableRawData().data()), cells,
              sizeof(vgt::TaggedObjectOccupancyCell) * static_cast<size_t>(nx * ny * nz));
  return map;
}

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
}  // namespace

extern "C"
{
// free_and_named == 0: ExtractSignedDistanceField<T>(objects_to_use, parameters)
// free_and_named != 0: ExtractFreeAndNamedObjectsSignedDistanceField<T>(parameters)
int vgt_ref_tagged_map_sdf(int scalar_bytes, const void* cells, int64_t nx, int64_t ny, int64_t nz,
                           double resolution, const uint32_t* object_ids, int64_t num_object_ids,
                           int unknown_is_filled, int add_virtual_border, int free_and_named,
                           void* sdf_out, void* min_max)
{
  try
  {
    const auto map = MakeTaggedMap(cells, nx, ny, nz, resolution);
    const std::vector<uint32_t> objects(object_ids, object_ids + num_object_ids);
    if (scalar_bytes == 8)
    {
      const auto parameters = Parameters<double>(unknown_is_filled, add_virtual_border);
      CopyOut(free_and_named != 0
                  ? map.ExtractFreeAndNamedObjectsSignedDistanceField<double>(parameters)
                  : map.ExtractSignedDistanceField<double>(objects, parameters),
              sdf_out, min_max);
    }
    else
    {
      const auto parameters = Parameters<float>(unknown_is_filled, add_virtual_border);
      CopyOut(free_and_named != 0
                  ? map.ExtractFreeAndNamedObjectsSignedDistanceField<float>(parameters)
                  : map.ExtractSignedDistanceField<float>(objects, parameters),
              sdf_out, min_max);
    }
    return 0;
  }
  catch (...)
  {
    return 1;
  }
}
}  // extern "C"
