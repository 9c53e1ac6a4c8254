// C++ parity test of the host adapter, written to read like the reference's own gtests
// (test/sdf_generation_test.cpp, test/pointcloud_voxelization_test.cpp) but without gtest, which
// this image does not have. It is compiled in the dev container against the REFERENCE'S
// pointcloud_voxelization_interface.hpp and vgt_namespace.hpp (from /root/reference/include) plus
// the stand-in third-party headers in oracle/ref_shim, linked with libvgt_b200.so, and runs on the
// GPU box (tests/test_gpu_cpp_adapter.py). Exit code 0 and "ADAPTER_TEST_OK" on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <vector>

#include <Eigen/Geometry>
#include <voxelized_geometry_tools/occupancy_map.hpp>
#include <voxelized_geometry_tools/pointcloud_voxelization_interface.hpp>
#include <voxelized_geometry_tools/signed_distance_field.hpp>

#include "b200_pointcloud_voxelization.hpp"
#include "b200_signed_distance_field_generation.hpp"

using namespace voxelized_geometry_tools;
using common_robotics_utilities::parallelism::DegreeOfParallelism;
using common_robotics_utilities::voxel_grid::GridIndex;
using common_robotics_utilities::voxel_grid::VoxelGridSizes;
namespace b200 = signed_distance_field_generation::b200;

static int g_failures = 0;
#define EXPECT_TRUE(cond)                                                        \
  do                                                                             \
  {                                                                              \
    if (!(cond))                                                                 \
    {                                                                            \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);              \
      g_failures++;                                                              \
    }                                                                            \
  } while (0)

constexpr double kExtremaTolerance = 0.0001;  // test/sdf_generation_test.cpp:22

template <typename ScalarType>
bool CloseEnough(const ScalarType a, const ScalarType b)
{
  return (a == b) || (std::abs(a - b) <= static_cast<ScalarType>(kExtremaTolerance));
}

template <typename ScalarType>
SignedDistanceFieldGenerationParameters<ScalarType> SDFGenerationParams()
{
  // test/sdf_generation_test.cpp:24-30
  return SignedDistanceFieldGenerationParameters<ScalarType>(
      std::numeric_limits<ScalarType>::infinity(), DegreeOfParallelism::None(), true, false);
}

OccupancyMap MakeMap(double resolution, double x_size, double y_size, double z_size, float fill)
{
  const auto grid_sizes =
      VoxelGridSizes::FromGridSizes(resolution, Eigen::Vector3d(x_size, y_size, z_size));
  return OccupancyMap(Eigen::Isometry3d::FromTranslation(-5.0, -5.0, -5.0), "test_frame",
                      grid_sizes, OccupancyCell(fill));
}

void FillBox(OccupancyMap& map, int64_t x0, int64_t x1, int64_t y0, int64_t y1, int64_t z0,
             int64_t z1)
{
  for (int64_t x = x0; x < x1; x++)
    for (int64_t y = y0; y < y1; y++)
      for (int64_t z = z0; z < z1; z++) map.SetIndex(x, y, z, OccupancyCell(1.0f));
}

// Both adapter paths must agree with each other and with the expectations.
template <typename ScalarType>
void TestSDFGeneration(const OccupancyMap& map, ScalarType expected_minimum,
                       ScalarType expected_maximum)
{
  const auto params = SDFGenerationParams<ScalarType>();
  const auto direct = b200::ExtractSignedDistanceFieldFromOccupancyMap(map, params);
  const std::function<bool(const GridIndex&)> is_filled_fn = [&](const GridIndex& index)
  {
    const float occupancy = map.GetIndexImmutable(index).Value().Occupancy();
    return (occupancy > 0.5) || (params.UnknownIsFilled() && occupancy == 0.5);
  };
  const auto via_predicate =
      b200::ExtractSignedDistanceField<OccupancyCell, std::vector<OccupancyCell>, ScalarType>(
          map, is_filled_fn, map.Frame(), params);
  for (const auto* sdf : {&direct, &via_predicate})
  {
    EXPECT_TRUE(sdf->IsLocked());
    EXPECT_TRUE(CloseEnough(sdf->GetMinimumMaximum().Minimum(), expected_minimum));
    EXPECT_TRUE(CloseEnough(sdf->GetMinimumMaximum().Maximum(), expected_maximum));
    for (int64_t x = 0; x < map.NumXVoxels(); x++)
      for (int64_t y = 0; y < map.NumYVoxels(); y++)
        for (int64_t z = 0; z < map.NumZVoxels(); z++)
        {
          const float occupancy = map.GetIndexImmutable(x, y, z).Value().Occupancy();
          const ScalarType value = sdf->GetIndexImmutable(x, y, z).Value();
          if (occupancy >= 0.5f) { EXPECT_TRUE(value < 0); }
          else { EXPECT_TRUE(value > 0); }
        }
  }
  EXPECT_TRUE(direct.GetImmutableRawData() == via_predicate.GetImmutableRawData());
}

void SdfTests()
{
  const float inf = std::numeric_limits<float>::infinity();
  // FullyFilledTest / FullyEmptyTest (:262-368)
  TestSDFGeneration<float>(MakeMap(0.25, 1.0, 2.0, 3.0, 1.0f), -inf, -inf);
  TestSDFGeneration<float>(MakeMap(0.25, 1.0, 2.0, 3.0, 0.0f), inf, inf);
  TestSDFGeneration<double>(MakeMap(0.25, 1.0, 2.0, 3.0, 0.0f),
                            std::numeric_limits<double>::infinity(),
                            std::numeric_limits<double>::infinity());
  {  // CenterObstacleTest (:370-443)
    OccupancyMap map = MakeMap(0.25, 1.0, 2.0, 3.0, 0.0f);
    FillBox(map, 1, 3, 2, 6, 3, 9);
    const double nominal = std::sqrt(0.25 * 0.25 + 0.5 * 0.5 + 0.75 * 0.75);
    TestSDFGeneration<float>(map, -0.25f, static_cast<float>(nominal));
    TestSDFGeneration<double>(map, -0.25, nominal);
  }
  {  // CornerObstacleTest (:445-513)
    OccupancyMap map = MakeMap(0.25, 1.0, 2.0, 3.0, 0.0f);
    FillBox(map, 0, 2, 0, 4, 0, 6);
    TestSDFGeneration<float>(map, -0.5f, 1.8708f);
    TestSDFGeneration<double>(map, -0.5, 1.8708);
  }
  {  // FaceObstacleTest (:515-585)
    OccupancyMap map = MakeMap(0.25, 1.0, 2.0, 3.0, 0.0f);
    FillBox(map, 0, map.NumXVoxels(), 0, map.NumYVoxels(), 0, 1);
    TestSDFGeneration<float>(map, -0.25f, 2.75f);
  }
  {  // PlanarExactTest (:704-903), every cell
    OccupancyMap map = MakeMap(1.0, 1.0, 4.0, 4.0, 0.0f);
    FillBox(map, 0, 1, 0, 2, 0, 2);
    const auto sdf =
        b200::ExtractSignedDistanceFieldFromOccupancyMap(map, SDFGenerationParams<float>());
    const float r2 = std::sqrt(2.0f), r5 = std::sqrt(5.0f), r8 = std::sqrt(8.0f);
    const float expected[4][4] = {{-2.0f, -1.0f, 1.0f, 2.0f},
                                  {-1.0f, -1.0f, 1.0f, 2.0f},
                                  {1.0f, 1.0f, r2, r5},
                                  {2.0f, 2.0f, r5, r8}};
    for (int64_t y = 0; y < 4; y++)
      for (int64_t z = 0; z < 4; z++)
        EXPECT_TRUE(sdf.GetIndexImmutable(0, y, z).Value() == expected[y][z]);
  }
  {  // CubeExactTest (:905-1056)
    OccupancyMap map = MakeMap(1.0, 2.0, 2.0, 2.0, 0.0f);
    FillBox(map, 0, 1, 0, 1, 0, 1);
    const auto sdf =
        b200::ExtractSignedDistanceFieldFromOccupancyMap(map, SDFGenerationParams<float>());
    EXPECT_TRUE(sdf.GetIndexImmutable(0, 0, 0).Value() == -1.0f);
    EXPECT_TRUE(sdf.GetIndexImmutable(0, 1, 1).Value() == std::sqrt(2.0f));
    EXPECT_TRUE(sdf.GetIndexImmutable(1, 1, 1).Value() == std::sqrt(3.0f));
  }
  // Locked result refuses mutation like the reference's (sdf.hpp:599-610).
  {
    OccupancyMap map = MakeMap(1.0, 1.0, 1.0, 4.0, 0.0f);
    FillBox(map, 0, 1, 0, 1, 0, 2);
    auto sdf = b200::ExtractSignedDistanceFieldFromOccupancyMap(map, SDFGenerationParams<float>());
    EXPECT_TRUE(!sdf.SetIndex(0, 0, 0, 5.0f));
    EXPECT_TRUE(sdf.GetIndexImmutable(0, 0, 0).Value() == -2.0f);
  }
}

// test/pointcloud_voxelization_test.cpp:31-82
class VectorVector3dPointCloudWrapper : public pointcloud_voxelization::PointCloudWrapper
{
public:
  void PushBack(double x, double y, double z)
  {
    points_.push_back(x);
    points_.push_back(y);
    points_.push_back(z);
  }
  double MaxRange() const override { return max_range_; }
  void SetMaxRange(const double max_range) override { max_range_ = max_range; }
  int64_t Size() const override { return static_cast<int64_t>(points_.size() / 3); }
  const Eigen::Isometry3d& PointCloudOriginTransform() const override { return origin_; }
  void SetPointCloudOriginTransform(const Eigen::Isometry3d& origin) override { origin_ = origin; }

private:
  void CopyPointLocationIntoDoublePtrImpl(const int64_t index, double* destination) const override
  {
    for (int i = 0; i < 3; i++) destination[i] = points_[static_cast<size_t>(index * 3 + i)];
  }
  void CopyPointLocationIntoFloatPtrImpl(const int64_t index, float* destination) const override
  {
    for (int i = 0; i < 3; i++)
      destination[i] = static_cast<float>(points_[static_cast<size_t>(index * 3 + i)]);
  }
  std::vector<double> points_;
  Eigen::Isometry3d origin_;
  double max_range_ = std::numeric_limits<double>::infinity();
};

Eigen::Isometry3d Rotation(const double r[3][3])
{
  Eigen::Isometry3d t;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t.at(i, j) = r[i][j];
  return t;
}

void VoxelizationTest()
{
  using namespace pointcloud_voxelization;
  // test/pointcloud_voxelization_test.cpp:160-246
  const auto grid_sizes = VoxelGridSizes::FromGridSizes(0.25, Eigen::Vector3d(2.0, 2.0, 2.0));
  OccupancyMap static_environment(Eigen::Isometry3d::FromTranslation(-1.0, -1.0, -1.0), "world",
                                  grid_sizes, OccupancyCell(0.0f));
  for (int64_t x = 0; x < 8; x++)
    for (int64_t y = 0; y < 8; y++) static_environment.SetIndex(x, y, 0, OccupancyCell(1.0f));
  const double co[3][3] = {{0, 0, 1}, {-1, 0, 0}, {0, -1, 0}};  // Rz(-pi/2) * Rx(-pi/2)
  const double rz90[3][3] = {{0, -1, 0}, {1, 0, 0}, {0, 0, 1}};
  const Eigen::Isometry3d X_CO = Rotation(co);
  auto cam1 = std::make_shared<VectorVector3dPointCloudWrapper>();
  cam1->SetPointCloudOriginTransform(Eigen::Isometry3d::FromTranslation(-2.0, 0.0, 0.0) * X_CO);
  auto cam2 = std::make_shared<VectorVector3dPointCloudWrapper>();
  cam2->SetPointCloudOriginTransform(Eigen::Isometry3d::FromTranslation(0.0, -2.0, 0.0)
                                     * Rotation(rz90) * X_CO);
  for (double x = -2.0; x <= 2.0; x += 0.03125)
    for (double y = -2.0; y <= 2.0; y += 0.03125)
    {
      cam1->PushBack(x, y, (x <= 0.0) ? 2.125 : 4.0);
      cam2->PushBack(x, y, (x >= 0.0) ? 2.125 : 4.0);
    }
  auto cam3 = std::make_shared<VectorVector3dPointCloudWrapper>();
  cam3->SetPointCloudOriginTransform(X_CO);
  const PointCloudVoxelizationFilterOptions filter_options(1.0, 1, 1);

  std::vector<std::string> log;
  const B200PointCloudVoxelizer voxelizer(
      {{"CUDA_DEVICE", 0}}, [&](const std::string& message) { log.push_back(message); });
  EXPECT_TRUE(!log.empty());

  const auto empty_voxelized =
      voxelizer.VoxelizePointClouds(static_environment, filter_options, {});
  for (int64_t x = 0; x < 8; x++)
    for (int64_t y = 0; y < 8; y++)
      for (int64_t z = 0; z < 8; z++)
      {
        const float occupancy = empty_voxelized.GetIndexImmutable(x, y, z).Value().Occupancy();
        EXPECT_TRUE(occupancy == ((z == 0) ? 1.0f : 0.5f));   // :84-111
      }

  bool runtime_logged = false;
  const auto voxelized = voxelizer.VoxelizePointClouds(
      static_environment, filter_options, {cam1, cam2, cam3},
      [&](const VoxelizerRuntime& runtime)
      { runtime_logged = runtime.RaycastingTime() >= 0.0 && runtime.FilteringTime() >= 0.0; });
  EXPECT_TRUE(runtime_logged);
  for (int64_t x = 0; x < 8; x++)
    for (int64_t y = 0; y < 8; y++)
      for (int64_t z = 0; z < 8; z++)
      {
        const float occupancy = voxelized.GetIndexImmutable(x, y, z).Value().Occupancy();
        if (z == 0) EXPECT_TRUE(occupancy == 1.0f);                         // :113-158
        if (x == 3 && y >= 3 && z >= 1) EXPECT_TRUE(occupancy == 0.0f);
        if (x >= 3 && y == 3 && z >= 1) EXPECT_TRUE(occupancy == 0.0f);
        if (x == 4 && y >= 4 && z >= 1) EXPECT_TRUE(occupancy == 1.0f);
        if (x >= 4 && y == 4 && z >= 1) EXPECT_TRUE(occupancy == 1.0f);
        if (x > 4 && y > 4 && z >= 1) EXPECT_TRUE(occupancy == 0.5f);
      }

  // null cloud -> invalid_argument (pcv_if.hpp:281-289); bad device -> runtime_error.
  bool threw = false;
  try { voxelizer.VoxelizePointClouds(static_environment, filter_options, {cam1, nullptr}); }
  catch (const std::invalid_argument&) { threw = true; }
  EXPECT_TRUE(threw);
  threw = false;
  try { B200PointCloudVoxelizer bad({{"CUDA_DEVICE", 99}}); }
  catch (const std::runtime_error&) { threw = true; }
  EXPECT_TRUE(threw);

  // voxelize -> SDF, the path end to end
  const auto sdf = b200::ExtractSignedDistanceFieldFromOccupancyMap(
      voxelized, SDFGenerationParams<float>());
  EXPECT_TRUE(sdf.GetIndexImmutable(0, 0, 0).Value() < 0.0f);
  EXPECT_TRUE(sdf.GetIndexImmutable(3, 3, 4).Value() > 0.0f);
}

int main()
{
  if (vgt_b200_device_count() < 1)
  {
    std::printf("no usable CUDA device: %s\n", vgt_b200_version());
    return 2;
  }
  SdfTests();
  VoxelizationTest();
  if (g_failures == 0)
  {
    std::printf("ADAPTER_TEST_OK\n");
    return 0;
  }
  std::printf("%d expectation(s) failed\n", g_failures);
  return 1;
}
