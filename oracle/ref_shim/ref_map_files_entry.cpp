// TEST INFRASTRUCTURE ONLY -- C entry points over the reference's own OccupancyMap file members
// (src/voxelized_geometry_tools/occupancy_map.cpp:100-193: Serialize, Deserialize, SaveToFile,
// LoadFromFile; derived members :56-84; the cell's encoding :23-46), compiled unmodified WITH THE
// REFERENCE'S OWN occupancy_map.hpp - this translation unit and occupancy_map.cpp are compiled
// with the reference's include directory in front of oracle/ref_shim, so the stand-in
// occupancy_map.hpp of the shim (which the other entry files use) is not seen here. They go into
// a library of their own (oracle/_ref/libvgt_ref_maps.so) to keep the two definitions apart.
// Also: OccupancyMap::ExtractSignedDistanceField<T> itself, the public entry this backend drops in
// for (tests/test_oracle_vs_reference.py, tests/test_gpu_sdf.py).
// Used by tests/test_grid_files.py. What this pins and what it cannot: see ref_grid_files_entry.cpp.
#include <cstdint>
#include <cstring>
#include <exception>
#include <limits>
#include <stdexcept>
#include <string>

#include <voxelized_geometry_tools/occupancy_map.hpp>

namespace
{
namespace vgt = voxelized_geometry_tools;
using common_robotics_utilities::voxel_grid::Vector3i64;
using common_robotics_utilities::voxel_grid::VoxelGridSizes;

int Report(char* message, int64_t capacity, const std::exception& error, int code)
{
  if (message != nullptr && capacity > 0)
  {
    std::strncpy(message, error.what(), static_cast<size_t>(capacity) - 1);
    message[capacity - 1] = 0;
  }
  return code;
}
}  // namespace

template <typename Scalar>
int ExtractFromMap(const float* occupancy, int64_t nx, int64_t ny, int64_t nz, double resolution,
                   int unknown_is_filled, int add_virtual_border, int threads, Scalar* sdf_out,
                   Scalar* min_max)
{
  using common_robotics_utilities::parallelism::DegreeOfParallelism;
  const auto sizes = VoxelGridSizes::FromVoxelCounts(resolution, Vector3i64(nx, ny, nz));
  vgt::OccupancyMap map(Eigen::Isometry3d::Identity(), "reference", sizes,
                        vgt::OccupancyCell(0.0f));
  std::memcpy(static_cast<void*>(map.GetMutableRawData().data()), occupancy,
              sizeof(float) * static_cast<size_t>(nx * ny * nz));
  const vgt::SignedDistanceFieldGenerationParameters<Scalar> parameters(
      std::numeric_limits<Scalar>::infinity(),
      threads <= 0 ? DegreeOfParallelism::FromOmp() : DegreeOfParallelism(threads),
      unknown_is_filled != 0, add_virtual_border != 0);
  // THE drop-in target: OccupancyMap::ExtractSignedDistanceField<T> (occupancy_map.hpp:174-216,
  // occupancy_map.cpp:250-260), predicate and all.
  const auto sdf = map.template ExtractSignedDistanceField<Scalar>(parameters);
  std::memcpy(sdf_out, sdf.GetImmutableRawData().data(),
              sizeof(Scalar) * static_cast<size_t>(nx * ny * nz));
  const auto extrema = sdf.GetMinimumMaximum();
  min_max[0] = extrema.Minimum();
  min_max[1] = extrema.Maximum();
  return 0;
}

extern "C"
{
// OccupancyMap::ExtractSignedDistanceFieldFloat / Double through the reference's own member.
int vgt_ref_map_extract_sdf(int scalar_bytes, const float* occupancy, int64_t nx, int64_t ny,
                            int64_t nz, double resolution, int unknown_is_filled,
                            int add_virtual_border, int threads, void* sdf_out, void* min_max)
{
  try
  {
    return scalar_bytes == 8
        ? ExtractFromMap(occupancy, nx, ny, nz, resolution, unknown_is_filled, add_virtual_border,
                         threads, static_cast<double*>(sdf_out), static_cast<double*>(min_max))
        : ExtractFromMap(occupancy, nx, ny, nz, resolution, unknown_is_filled, add_virtual_border,
                         threads, static_cast<float*>(sdf_out), static_cast<float*>(min_max));
  }
  catch (...)
  {
    return 1;
  }
}

// 0 = ok, 1 = std::invalid_argument, 2 = any other exception, 3 = caller buffer too small
int vgt_ref_map_save_to_file(const float* occupancy, int64_t nx, int64_t ny, int64_t nz,
                             double resolution, const double* origin_column_major,
                             const char* frame, float default_occupancy, float oob_occupancy,
                             const char* path, int compress, char* message,
                             int64_t message_capacity)
{
  try
  {
    Eigen::Isometry3d origin;
    std::memcpy(origin.data(), origin_column_major, sizeof(double) * 16);
    const auto sizes = VoxelGridSizes::FromVoxelCounts(resolution, Vector3i64(nx, ny, nz));
    vgt::OccupancyMap map(origin, frame, sizes, vgt::OccupancyCell(default_occupancy),
                          vgt::OccupancyCell(oob_occupancy));
    static_assert(sizeof(vgt::OccupancyCell) == sizeof(float), "packed cells");
    std::memcpy(static_cast<void*>(map.GetMutableRawData().data()), occupancy,
                sizeof(float) * static_cast<size_t>(nx * ny * nz));
    vgt::OccupancyMap::SaveToFile(map, path, compress != 0);
    return 0;
  }
  catch (const std::invalid_argument& error) { return Report(message, message_capacity, error, 1); }
  catch (const std::exception& error) { return Report(message, message_capacity, error, 2); }
}

int vgt_ref_map_load_from_file(const char* path, float* occupancy, int64_t capacity, int64_t* dims,
                               double* resolution, double* origin_column_major, char* frame,
                               int64_t frame_capacity, float* default_and_oob, char* message,
                               int64_t message_capacity)
{
  try
  {
    const vgt::OccupancyMap map = vgt::OccupancyMap::LoadFromFile(path);
    dims[0] = map.NumXVoxels();
    dims[1] = map.NumYVoxels();
    dims[2] = map.NumZVoxels();
    *resolution = map.Resolution();
    std::memcpy(origin_column_major, map.OriginTransform().data(), sizeof(double) * 16);
    std::strncpy(frame, map.Frame().c_str(), static_cast<size_t>(frame_capacity) - 1);
    frame[frame_capacity - 1] = 0;
    default_and_oob[0] = map.DefaultValue().Occupancy();
    default_and_oob[1] = map.OOBValue().Occupancy();
    if (map.NumTotalVoxels() > capacity)
    {
      return 3;
    }
    std::memcpy(occupancy, static_cast<const void*>(map.GetImmutableRawData().data()),
                sizeof(float) * static_cast<size_t>(map.NumTotalVoxels()));
    return 0;
  }
  catch (const std::invalid_argument& error) { return Report(message, message_capacity, error, 1); }
  catch (const std::exception& error) { return Report(message, message_capacity, error, 2); }
}
}  // extern "C"
