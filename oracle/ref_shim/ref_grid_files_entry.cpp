// TEST INFRASTRUCTURE ONLY -- C entry points over the reference's own file members of
// SignedDistanceField<T> (include/voxelized_geometry_tools/signed_distance_field.hpp:622-722:
// Serialize, Deserialize, SaveToFile, LoadFromFile, and DerivedSerializeSelf / DerivedDeserializeSelf
// at :551-596), compiled unmodified over the stand-in third-party headers of oracle/ref_shim.
// Used by tests/test_grid_files.py to check csrc/grid_files.cu and oracle/grid_files_oracle.py.
//
// What this pins: the statements of the reference header - the "SDFZ" / "SDFR" magics, what the
// compress switch does, the derived members (frame string, then the locked byte) and their
// place after the grid's own bytes, Lock() on load. What it cannot pin: the grid's own bytes and
// the primitive encodings come from common_robotics_utilities, which is not in the reference
// tree; ref_shim restates them (serialization.hpp, voxel_grid.hpp, zlib_helpers.hpp).
#include <cstdint>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>

#include <voxelized_geometry_tools/signed_distance_field.hpp>

namespace
{
namespace vgt = voxelized_geometry_tools;
using common_robotics_utilities::voxel_grid::Vector3i64;
using common_robotics_utilities::voxel_grid::VoxelGridSizes;

template <typename Scalar>
int Save(const Scalar* values, int64_t nx, int64_t ny, int64_t nz, double resolution,
         const double* origin_column_major, const char* frame, int locked, double oob_value,
         const char* path, int compress)
{
  Eigen::Isometry3d origin;
  std::memcpy(origin.data(), origin_column_major, sizeof(double) * 16);
  const auto sizes = VoxelGridSizes::FromVoxelCounts(resolution, Vector3i64(nx, ny, nz));
  vgt::SignedDistanceField<Scalar> field(origin, frame, sizes, static_cast<Scalar>(oob_value));
  std::memcpy(field.GetMutableRawData().data(), values,
              sizeof(Scalar) * static_cast<size_t>(nx * ny * nz));
  if (locked != 0)
  {
    field.Lock();
  }
  vgt::SignedDistanceField<Scalar>::SaveToFile(field, path, compress != 0);
  return 0;
}

template <typename Scalar>
int Load(const char* path, Scalar* values, int64_t capacity, int64_t* dims, double* resolution,
         double* origin_column_major, char* frame, int64_t frame_capacity, int* locked,
         double* default_and_oob, double* min_max)
{
  const auto field = vgt::SignedDistanceField<Scalar>::LoadFromFile(path);
  dims[0] = field.NumXVoxels();
  dims[1] = field.NumYVoxels();
  dims[2] = field.NumZVoxels();
  *resolution = field.VoxelXSize();
  std::memcpy(origin_column_major, field.OriginTransform().data(), sizeof(double) * 16);
  std::strncpy(frame, field.Frame().c_str(), static_cast<size_t>(frame_capacity) - 1);
  frame[frame_capacity - 1] = 0;
  *locked = field.IsLocked() ? 1 : 0;
  default_and_oob[0] = static_cast<double>(field.DefaultValue());
  default_and_oob[1] = static_cast<double>(field.OOBValue());
  if (field.NumTotalVoxels() > capacity)
  {
    return 3;
  }
  std::memcpy(values, field.GetImmutableRawData().data(),
              sizeof(Scalar) * static_cast<size_t>(field.NumTotalVoxels()));
  if (field.IsLocked())
  {
    const auto extrema = field.GetMinimumMaximum();
    min_max[0] = static_cast<double>(extrema.Minimum());
    min_max[1] = static_cast<double>(extrema.Maximum());
  }
  return 0;
}

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

extern "C"
{
// 0 = ok, 1 = std::invalid_argument, 2 = any other exception, 3 = caller buffer too small
int vgt_ref_sdf_save_to_file(int scalar_bytes, const void* values, int64_t nx, int64_t ny,
                             int64_t nz, double resolution, const double* origin_column_major,
                             const char* frame, int locked, double oob_value, const char* path,
                             int compress, char* message, int64_t message_capacity)
{
  try
  {
    return scalar_bytes == 8
        ? Save(static_cast<const double*>(values), nx, ny, nz, resolution, origin_column_major,
               frame, locked, oob_value, path, compress)
        : Save(static_cast<const float*>(values), nx, ny, nz, resolution, origin_column_major,
               frame, locked, oob_value, path, compress);
  }
  catch (const std::invalid_argument& error) { return Report(message, message_capacity, error, 1); }
  catch (const std::exception& error) { return Report(message, message_capacity, error, 2); }
}

int vgt_ref_sdf_load_from_file(int scalar_bytes, const char* path, void* values, int64_t capacity,
                               int64_t* dims, double* resolution, double* origin_column_major,
                               char* frame, int64_t frame_capacity, int* locked,
                               double* default_and_oob, double* min_max, char* message,
                               int64_t message_capacity)
{
  try
  {
    return scalar_bytes == 8
        ? Load(path, static_cast<double*>(values), capacity, dims, resolution,
               origin_column_major, frame, frame_capacity, locked, default_and_oob, min_max)
        : Load(path, static_cast<float*>(values), capacity, dims, resolution,
               origin_column_major, frame, frame_capacity, locked, default_and_oob, min_max);
  }
  catch (const std::invalid_argument& error) { return Report(message, message_capacity, error, 1); }
  catch (const std::exception& error) { return Report(message, message_capacity, error, 2); }
}
}  // extern "C"
