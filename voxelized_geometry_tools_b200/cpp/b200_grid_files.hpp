// C++ host adapter: the reference's SDF files through libvgt_b200 (vgt_b200_grid_file_*,
// csrc/grid_files.cu), with the names and argument meaning of
//   SignedDistanceField<T>::SaveToFile / LoadFromFile
//   (include/voxelized_geometry_tools/signed_distance_field.hpp:643-722).
// A host that has the reference's class keeps using its own members; these functions are for
// a field that lives on the device (SaveDeviceSignedDistanceFieldToFile) and for checking, in
// one binary, that both write the same bytes (cpp/test/adapter_test.cpp GridFileTest).
#pragma once

#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include <voxelized_geometry_tools/signed_distance_field.hpp>

#include "vgt_b200.h"

namespace voxelized_geometry_tools
{
VGT_NAMESPACE_BEGIN
namespace grid_files
{
namespace b200
{
namespace grid_file_internal
{
inline void Check(const int code)
{
  if (code == VGT_B200_OK)
  {
    return;
  }
  if (code == VGT_B200_ERR_INVALID_ARGUMENT)
  {
    throw std::invalid_argument(vgt_b200_last_error());  // "File does not exist", ...
  }
  throw std::runtime_error(vgt_b200_last_error());
}

template <typename ScalarType>
constexpr int Kind()
{
  static_assert(std::is_same<ScalarType, float>::value || std::is_same<ScalarType, double>::value,
                "SignedDistanceField<float> or <double>");
  return std::is_same<ScalarType, float>::value ? VGT_B200_GRID_FILE_SDF_F32
                                                : VGT_B200_GRID_FILE_SDF_F64;
}

template <typename ScalarType>
vgt_b200_grid_file_info Describe(const SignedDistanceField<ScalarType>& sdf)
{
  vgt_b200_grid_file_info info;
  std::memset(&info, 0, sizeof(info));
  info.nx = sdf.NumXVoxels();
  info.ny = sdf.NumYVoxels();
  info.nz = sdf.NumZVoxels();
  const auto voxel_sizes = sdf.VoxelSizes();
  for (int i = 0; i < 3; i++)
  {
    info.voxel_size[i] = voxel_sizes(i);
  }
  std::memcpy(info.origin_transform, sdf.OriginTransform().data(), sizeof(double) * 16);
  info.default_value = static_cast<double>(sdf.DefaultValue());
  info.oob_value = static_cast<double>(sdf.OOBValue());
  info.initialized = sdf.IsInitialized() ? 1 : 0;
  info.locked = sdf.IsLocked() ? 1 : 0;
  return info;
}
}  // namespace grid_file_internal

// SignedDistanceField<T>::SaveToFile (signed_distance_field.hpp:643-668).
template <typename ScalarType>
void SaveSignedDistanceFieldToFile(const SignedDistanceField<ScalarType>& sdf,
                                   const std::string& filepath, const bool compress)
{
  const vgt_b200_grid_file_info info = grid_file_internal::Describe(sdf);
  grid_file_internal::Check(vgt_b200_grid_file_save(
      filepath.c_str(), grid_file_internal::Kind<ScalarType>(), compress ? 1 : 0,
      sdf.GetImmutableRawData().data(), &info, sdf.Frame().c_str()));
}

// The same for values that live on `device` (d_values = ScalarType[nx * ny * nz], x slowest);
// `like` supplies sizes, origin transform, frame and out-of-bounds value.
template <typename ScalarType>
void SaveDeviceSignedDistanceFieldToFile(const ScalarType* d_values,
                                         const SignedDistanceField<ScalarType>& like,
                                         const bool locked, const std::string& filepath,
                                         const bool compress, const int device = 0,
                                         void* stream = nullptr)
{
  vgt_b200_grid_file_info info = grid_file_internal::Describe(like);
  info.locked = locked ? 1 : 0;
  grid_file_internal::Check(vgt_b200_grid_file_save_dev(
      filepath.c_str(), grid_file_internal::Kind<ScalarType>(), compress ? 1 : 0, d_values, &info,
      like.Frame().c_str(), device, stream));
}

// SignedDistanceField<T>::LoadFromFile (signed_distance_field.hpp:670-722): a field saved locked
// comes back locked (Lock() recomputes the extrema, :579-592).
template <typename ScalarType>
SignedDistanceField<ScalarType> LoadSignedDistanceFieldFromFile(const std::string& filepath)
{
  using common_robotics_utilities::voxel_grid::Vector3i64;
  using common_robotics_utilities::voxel_grid::VoxelGridSizes;
  vgt_b200_grid_file_info info;
  grid_file_internal::Check(vgt_b200_grid_file_probe(
      filepath.c_str(), grid_file_internal::Kind<ScalarType>(), &info, nullptr, 0));
  std::vector<char> frame(static_cast<size_t>(info.frame_length) + 1, 0);
  const auto sizes = VoxelGridSizes::FromVoxelCounts(
      info.voxel_size[0], Vector3i64(info.nx, info.ny, info.nz));
  Eigen::Isometry3d origin;
  std::memcpy(origin.data(), info.origin_transform, sizeof(double) * 16);
  SignedDistanceField<ScalarType> sdf(origin, "", sizes, static_cast<ScalarType>(info.oob_value));
  grid_file_internal::Check(vgt_b200_grid_file_load(
      filepath.c_str(), grid_file_internal::Kind<ScalarType>(), sdf.GetMutableRawData().data(),
      info.nx * info.ny * info.nz, &info, frame.data(), static_cast<int64_t>(frame.size())));
  sdf.SetFrame(std::string(frame.data()));
  if (info.locked != 0)
  {
    sdf.Lock();
  }
  return sdf;
}
}  // namespace b200
}  // namespace grid_files
VGT_NAMESPACE_END
}  // namespace voxelized_geometry_tools
