// TEST INFRASTRUCTURE ONLY -- C entry points over the REFERENCE'S OWN SDF generation code.
//
// This translation unit includes the reference's signed_distance_field_generation.hpp unmodified
// and is linked with the reference's signed_distance_field_generation.cpp compiled unmodified from
// /root/reference (see oracle/Makefile, target `ref`). Only the third-party layer underneath
// (Eigen, common_robotics_utilities) and the SDF container are stand-ins (oracle/ref_shim/...).
// So the 1-D transforms, the pass structure, the field marking / combine loops and the virtual
// border logic that run here are literally the reference's.
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>

#include <voxelized_geometry_tools/signed_distance_field_generation.hpp>

namespace
{
namespace vgt = voxelized_geometry_tools;
namespace sdfgen = voxelized_geometry_tools::signed_distance_field_generation::internal;
using common_robotics_utilities::parallelism::DegreeOfParallelism;
using common_robotics_utilities::voxel_grid::GridIndex;
using common_robotics_utilities::voxel_grid::Vector3i64;
using common_robotics_utilities::voxel_grid::VoxelGrid;
using common_robotics_utilities::voxel_grid::VoxelGridSizes;

DegreeOfParallelism Threads(int threads)
{
  if (threads <= 0) { return DegreeOfParallelism::FromOmp(); }
  return DegreeOfParallelism(threads);
}

template <typename Scalar>
int Extract(const float* occupancy, int64_t nx, int64_t ny, int64_t nz, double resolution,
            int unknown_is_filled, int add_virtual_border, int threads, Scalar* sdf_out,
            Scalar* min_max)
{
  const auto sizes = VoxelGridSizes::FromVoxelCounts(resolution, Vector3i64(nx, ny, nz));
  VoxelGrid<float> grid(Eigen::Isometry3d::Identity(), sizes, 0.0f);
  std::memcpy(grid.GetMutableRawData().data(), occupancy,
              sizeof(float) * static_cast<size_t>(nx * ny * nz));
  // The predicate OccupancyMap::ExtractSignedDistanceField builds
  // (include/voxelized_geometry_tools/occupancy_map.hpp:181-205).
  const std::function<bool(const GridIndex&)> is_filled_fn = [&](const GridIndex& index)
  {
    const auto query = grid.GetIndexImmutable(index);
    if (query)
    {
      const float occ = query.Value();
      if (occ > 0.5) { return true; }
      if (unknown_is_filled && (occ == 0.5)) { return true; }
      return false;
    }
    throw std::runtime_error("index out of grid bounds");
  };
  const vgt::SignedDistanceFieldGenerationParameters<Scalar> parameters(
      std::numeric_limits<Scalar>::infinity(), Threads(threads), unknown_is_filled != 0,
      add_virtual_border != 0);
  const auto sdf = sdfgen::ExtractSignedDistanceField<float, std::vector<float>, Scalar>(
      grid, is_filled_fn, "reference", parameters);
  std::memcpy(sdf_out, sdf.GetImmutableRawData().data(),
              sizeof(Scalar) * static_cast<size_t>(nx * ny * nz));
  if (min_max != nullptr)
  {
    const auto extrema = sdf.GetMinimumMaximum();
    min_max[0] = extrema.Minimum();
    min_max[1] = extrema.Maximum();
  }
  return 0;
}
}  // namespace

extern "C"
{
int vgt_ref_sdf_f32(const float* occupancy, int64_t nx, int64_t ny, int64_t nz, double resolution,
                    int unknown_is_filled, int add_virtual_border, int threads, float* sdf,
                    float* min_max)
{
  try
  {
    return Extract<float>(occupancy, nx, ny, nz, resolution, unknown_is_filled,
                          add_virtual_border, threads, sdf, min_max);
  }
  catch (...)
  {
    return 1;
  }
}

int vgt_ref_sdf_f64(const float* occupancy, int64_t nx, int64_t ny, int64_t nz, double resolution,
                    int unknown_is_filled, int add_virtual_border, int threads, double* sdf,
                    double* min_max)
{
  try
  {
    return Extract<double>(occupancy, nx, ny, nz, resolution, unknown_is_filled,
                           add_virtual_border, threads, sdf, min_max);
  }
  catch (...)
  {
    return 1;
  }
}

// ComputeDistanceFieldTransformInPlace on a caller-provided double field (sdfgen.hpp:34-37).
int vgt_ref_transform_inplace_f64(double* field, int64_t nx, int64_t ny, int64_t nz, int threads)
{
  try
  {
    const auto sizes = VoxelGridSizes::FromVoxelCounts(1.0, Vector3i64(nx, ny, nz));
    sdfgen::EDTDistanceField grid(Eigen::Isometry3d::Identity(), sizes, 0.0);
    const size_t bytes = sizeof(double) * static_cast<size_t>(nx * ny * nz);
    std::memcpy(grid.GetMutableRawData().data(), field, bytes);
    sdfgen::ComputeDistanceFieldTransformInPlace(Threads(threads), grid);
    std::memcpy(field, grid.GetImmutableRawData().data(), bytes);
    return 0;
  }
  catch (...)
  {
    return 1;
  }
}
}  // extern "C"
