// TEST INFRASTRUCTURE ONLY -- C entry points over the reference's own SignedDistanceField class
// (include/voxelized_geometry_tools/signed_distance_field.hpp, compiled unmodified over the
// stand-in third-party headers of oracle/ref_shim): the query members the device kernels of
// csrc/sdf_queries.cu replace, called one point at a time exactly as a user of the class would.
// Used by tests/test_oracle_vs_reference.py to pin oracle/sdf_queries_oracle.py.
//
// What this pins and what it cannot: every statement of the reference header itself (index
// selection, corrected centre distances, gradients, the projection loop, the extrema walk) runs
// here as written. The trilinear blend is common_robotics_utilities::math::TrilinearInterpolate,
// which is not in the reference tree; ref_shim/common_robotics_utilities/math.hpp restates it,
// so EstimateLocationDistance (and the projection, which calls it) stay "blend unpinned".
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>

#include <voxelized_geometry_tools/signed_distance_field.hpp>

namespace
{
namespace vgt = voxelized_geometry_tools;
using common_robotics_utilities::voxel_grid::Vector3i64;
using common_robotics_utilities::voxel_grid::VoxelGridSizes;

constexpr uint8_t kNoValue = 0;
constexpr uint8_t kValue = 1;
constexpr uint8_t kThrows = 2;

vgt::SignedDistanceField<float> MakeField(const float* sdf, int64_t nx, int64_t ny, int64_t nz,
                                          double resolution, const double* origin_column_major)
{
  Eigen::Isometry3d origin;
  std::memcpy(origin.data(), origin_column_major, sizeof(double) * 16);
  const auto sizes = VoxelGridSizes::FromVoxelCounts(resolution, Vector3i64(nx, ny, nz));
  vgt::SignedDistanceField<float> field(origin, "reference", sizes, 0.0f);
  std::memcpy(field.GetMutableRawData().data(), sdf,
              sizeof(float) * static_cast<size_t>(nx * ny * nz));
  return field;
}
}  // namespace

extern "C"
{
// kind 0: EstimateLocationDistance4d          -> 1 value per point
// kind 1: GetLocationCoarseGradient4d (a = enable_edge_gradients)      -> 3 values
// kind 2: GetLocationFineGradient4d (a = nominal_window_size)          -> 3 values
// kind 3: ProjectLocationOutOfCollisionToMinimumDistance4d (a = minimum_distance,
//         b = stepsize_multiplier)                                      -> 3 values
int vgt_ref_sdf_query(const float* sdf, int64_t nx, int64_t ny, int64_t nz, double resolution,
                      const double* origin_column_major, int kind, double a, double b,
                      const double* points, int64_t count, double* values, uint8_t* status)
{
  try
  {
    const auto field = MakeField(sdf, nx, ny, nz, resolution, origin_column_major);
    const int width = kind == 0 ? 1 : 3;
    for (int64_t i = 0; i < count; i++)
    {
      const Eigen::Vector4d location(points[3 * i], points[3 * i + 1], points[3 * i + 2], 1.0);
      double* out = values + width * i;
      for (int k = 0; k < width; k++) { out[k] = 0.0; }
      try
      {
        if (kind == 0)
        {
          const auto query = field.EstimateLocationDistance4d(location);
          status[i] = query.HasValue() ? kValue : kNoValue;
          if (query.HasValue()) { out[0] = query.Value(); }
        }
        else if (kind == 1 || kind == 2)
        {
          const auto query = kind == 1 ? field.GetLocationCoarseGradient4d(location, a != 0.0)
                                       : field.GetLocationFineGradient4d(location, a);
          status[i] = query.HasValue() ? kValue : kNoValue;
          if (query.HasValue())
          {
            for (int k = 0; k < 3; k++) { out[k] = query.Value()(k); }
          }
        }
        else
        {
          const auto query
              = field.ProjectLocationOutOfCollisionToMinimumDistance4d(location, a, b);
          status[i] = query.HasValue() ? kValue : kNoValue;
          if (query.HasValue())
          {
            for (int k = 0; k < 3; k++) { out[k] = query.Value()(k); }
          }
        }
      }
      catch (const std::exception&)
      {
        status[i] = kThrows;
      }
    }
    return 0;
  }
  catch (const std::exception&)
  {
    return 1;
  }
}

// ComputeLocalExtremaMap: double[nx*ny*nz*3].
int vgt_ref_sdf_local_extrema_map(const float* sdf, int64_t nx, int64_t ny, int64_t nz,
                                  double resolution, const double* origin_column_major,
                                  double* extrema)
{
  try
  {
    const auto field = MakeField(sdf, nx, ny, nz, resolution, origin_column_major);
    const auto map = field.ComputeLocalExtremaMap();
    const auto& cells = map.GetImmutableRawData();
    for (size_t i = 0; i < cells.size(); i++)
    {
      extrema[3 * i] = cells[i](0);
      extrema[3 * i + 1] = cells[i](1);
      extrema[3 * i + 2] = cells[i](2);
    }
    return 0;
  }
  catch (const std::exception&)
  {
    return 1;
  }
}
}  // extern "C"
