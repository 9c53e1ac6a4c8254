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

#include <chrono>
#include <cstring>
#include <set>
#include <string>

#include <Eigen/Geometry>
#include <voxelized_geometry_tools/mesh_rasterizer.hpp>
#include <voxelized_geometry_tools/occupancy_component_map.hpp>
#include <voxelized_geometry_tools/occupancy_map.hpp>
#include <voxelized_geometry_tools/pointcloud_voxelization.hpp>
#include <voxelized_geometry_tools/pointcloud_voxelization_interface.hpp>
#include <voxelized_geometry_tools/signed_distance_field.hpp>
#include <voxelized_geometry_tools/tagged_object_occupancy_component_map.hpp>
#include <voxelized_geometry_tools/tagged_object_occupancy_map.hpp>

#include "b200_cell_map_signed_distance_fields.hpp"
#include "b200_grid_files.hpp"
#include "b200_mesh_rasterizer.hpp"
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

// With two or more devices: the multi-device overload gives the one-device result bit for bit.
void MultiDeviceTest()
{
  const int devices = vgt_b200_device_count();
  if (devices < 2)
  {
    return;
  }
  OccupancyMap map = MakeMap(0.25, 8.0, 6.0, 7.0, 0.0f);
  FillBox(map, 3, 9, 2, 11, 5, 20);
  FillBox(map, 20, 31, 14, 19, 1, 6);
  for (const bool border : {false, true})
  {
    const SignedDistanceFieldGenerationParameters<float> params(
        std::numeric_limits<float>::infinity(), DegreeOfParallelism::None(), true, border);
    const auto single = b200::ExtractSignedDistanceFieldFromOccupancyMap(map, params);
    std::vector<int> listed;
    for (int d = 0; d < devices; d++) { listed.push_back(d); }
    const auto multi = b200::ExtractSignedDistanceFieldFromOccupancyMap(map, params, listed);
    EXPECT_TRUE(multi.IsLocked());
    EXPECT_TRUE(multi.GetImmutableRawData() == single.GetImmutableRawData());
    EXPECT_TRUE(multi.GetMinimumMaximum().Minimum() == single.GetMinimumMaximum().Minimum());
    EXPECT_TRUE(multi.GetMinimumMaximum().Maximum() == single.GetMinimumMaximum().Maximum());
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

// A small deterministic scene for the cell maps: occupancy in {0, 0.5, 1}, object ids 0..3.
struct CellScene
{
  int64_t nx = 9, ny = 12, nz = 10;
  double resolution = 0.25;
  float Occupancy(int64_t x, int64_t y, int64_t z) const
  {
    const uint32_t h = static_cast<uint32_t>(x * 73856093u) ^ static_cast<uint32_t>(y * 19349663u)
        ^ static_cast<uint32_t>(z * 83492791u);
    if ((x / 3 + y / 4 + z / 3) % 3 == 0) { return 1.0f; }
    return (h % 17u == 0u) ? 0.5f : 0.0f;
  }
  uint32_t ObjectId(int64_t x, int64_t y, int64_t z) const
  {
    return static_cast<uint32_t>((x / 3 + 2 * (y / 4) + z / 5) % 4);
  }
  VoxelGridSizes Sizes() const
  {
    return VoxelGridSizes::FromVoxelCounts(
        resolution, common_robotics_utilities::voxel_grid::Vector3i64(nx, ny, nz));
  }
};

template <typename ScalarType>
bool SameValues(const SignedDistanceField<ScalarType>& a, const SignedDistanceField<ScalarType>& b)
{
  return a.GetImmutableRawData() == b.GetImmutableRawData()
      && a.GetMinimumMaximum().Minimum() == b.GetMinimumMaximum().Minimum()
      && a.GetMinimumMaximum().Maximum() == b.GetMinimumMaximum().Maximum() && a.IsLocked()
      && b.IsLocked();
}

// The three cell-map types: same SDF as OccupancyMap for the same occupancy
// (test/sdf_generation_test.cpp:152-189), object lists, per-object batches, free-and-named merge.
template <typename ScalarType>
void CellMapTests()
{
  const CellScene scene;
  const auto origin = Eigen::Isometry3d::FromTranslation(1.0, 2.0, 3.0);
  OccupancyMap occupancy_map(origin, "cells", scene.Sizes(), OccupancyCell(0.0f));
  OccupancyComponentMap component_map(origin, "cells", scene.Sizes(), OccupancyComponentCell(0.0f));
  TaggedObjectOccupancyMap tagged_map(origin, "cells", scene.Sizes(),
                                      TaggedObjectOccupancyCell(0.0f));
  TaggedObjectOccupancyComponentMap tagged_component_map(
      origin, "cells", scene.Sizes(), TaggedObjectOccupancyComponentCell(0.0f));
  for (int64_t x = 0; x < scene.nx; x++)
    for (int64_t y = 0; y < scene.ny; y++)
      for (int64_t z = 0; z < scene.nz; z++)
      {
        const float occupancy = scene.Occupancy(x, y, z);
        const uint32_t id = scene.ObjectId(x, y, z);
        occupancy_map.SetIndex(x, y, z, OccupancyCell(occupancy));
        component_map.SetIndex(x, y, z, OccupancyComponentCell(occupancy, 7u));
        tagged_map.SetIndex(x, y, z, TaggedObjectOccupancyCell(occupancy, id));
        tagged_component_map.SetIndex(
            x, y, z, TaggedObjectOccupancyComponentCell(occupancy, id, 5u, 9u));
      }
  for (const bool unknown_is_filled : {true, false})
    for (const bool border : {false, true})
    {
      const SignedDistanceFieldGenerationParameters<ScalarType> params(
          std::numeric_limits<ScalarType>::infinity(), DegreeOfParallelism::None(),
          unknown_is_filled, border);
      const auto want = b200::ExtractSignedDistanceFieldFromOccupancyMap(occupancy_map, params);
      EXPECT_TRUE(SameValues(want, b200::ExtractSignedDistanceFieldFromCellMap(
                                       component_map, {}, params)));
      EXPECT_TRUE(SameValues(want, b200::ExtractSignedDistanceFieldFromCellMap(
                                       tagged_map, {}, params)));
      EXPECT_TRUE(SameValues(want, b200::ExtractSignedDistanceFieldFromCellMap(
                                       tagged_component_map, {}, params)));
      // objects_to_use (tagged_object_occupancy_map.hpp:199-247) against the opaque-predicate
      // path with the reference's own predicate
      const std::vector<uint32_t> objects_to_use = {3u, 1u, 3u};
      const std::set<uint32_t> wanted(objects_to_use.begin(), objects_to_use.end());
      const std::function<bool(const GridIndex&)> listed_fn = [&](const GridIndex& index)
      {
        const auto cell = tagged_map.GetIndexImmutable(index).Value();
        if (wanted.count(cell.ObjectId()) != 1) { return false; }
        return (cell.Occupancy() > 0.5) || (unknown_is_filled && cell.Occupancy() == 0.5);
      };
      const auto listed_want = b200::ExtractSignedDistanceField<
          TaggedObjectOccupancyCell, std::vector<TaggedObjectOccupancyCell>, ScalarType>(
          tagged_map, listed_fn, tagged_map.Frame(), params);
      EXPECT_TRUE(SameValues(listed_want, b200::ExtractSignedDistanceFieldFromCellMap(
                                              tagged_map, objects_to_use, params)));
      EXPECT_TRUE(SameValues(listed_want, b200::ExtractSignedDistanceFieldFromCellMap(
                                              tagged_component_map, objects_to_use, params)));
      // MakeSeparateObjectSDFs / MakeAllObjectSDFs (:249-291)
      const auto separate = b200::MakeSeparateObjectSDFs(tagged_map, {2u, 1u}, params);
      EXPECT_TRUE(separate.size() == 2);
      for (const auto& id_and_sdf : separate)
      {
        EXPECT_TRUE(SameValues(id_and_sdf.second, b200::ExtractSignedDistanceFieldFromCellMap(
                                                      tagged_map, {id_and_sdf.first}, params)));
      }
      const auto all_objects = b200::MakeAllObjectSDFs(tagged_component_map, params);
      EXPECT_TRUE(all_objects.size() == 3 && all_objects.count(0u) == 0);
      EXPECT_TRUE(SameValues(all_objects.at(2u), separate.at(2u)));
      // ExtractFreeAndNamedObjectsSignedDistanceField (:293-378): free >= 0 -> free; else
      // named <= -0 -> named; else 0
      const std::function<bool(const GridIndex&)> named_fn = [&](const GridIndex& index)
      {
        const auto cell = tagged_map.GetIndexImmutable(index).Value();
        if (cell.ObjectId() == 0u) { return false; }
        return (cell.Occupancy() > 0.5) || (unknown_is_filled && cell.Occupancy() == 0.5);
      };
      const auto named = b200::ExtractSignedDistanceField<
          TaggedObjectOccupancyCell, std::vector<TaggedObjectOccupancyCell>, ScalarType>(
          tagged_map, named_fn, tagged_map.Frame(), params);
      const auto merged = b200::ExtractFreeAndNamedObjectsSignedDistanceField(tagged_map, params);
      EXPECT_TRUE(merged.IsLocked());
      bool merge_ok = true;
      for (size_t i = 0; i < merged.GetImmutableRawData().size(); i++)
      {
        const ScalarType free_value = want.GetImmutableRawData()[i];
        const ScalarType named_value = named.GetImmutableRawData()[i];
        const ScalarType expected =
            (free_value >= 0) ? free_value : ((named_value <= -0.0) ? named_value : ScalarType(0));
        merge_ok = merge_ok && (merged.GetImmutableRawData()[i] == expected);
      }
      EXPECT_TRUE(merge_ok);
      EXPECT_TRUE(SameValues(merged, b200::ExtractFreeAndNamedObjectsSignedDistanceField(
                                         tagged_component_map, params)));
    }
}

// b200::ComputeDistanceFieldTransformInPlace against a brute-force transform.
void TransformTest()
{
  using common_robotics_utilities::voxel_grid::VoxelGrid;
  const int64_t nx = 7, ny = 9, nz = 11;
  const double inf = std::numeric_limits<double>::infinity();
  VoxelGrid<double> field(
      Eigen::Isometry3d::Identity(),
      VoxelGridSizes::FromVoxelCounts(1.0, common_robotics_utilities::voxel_grid::Vector3i64(nx, ny, nz)),
      inf);
  std::vector<double> samples(static_cast<size_t>(nx * ny * nz), inf);
  for (int64_t x = 0; x < nx; x++)
    for (int64_t y = 0; y < ny; y++)
      for (int64_t z = 0; z < nz; z++)
      {
        const uint32_t h = static_cast<uint32_t>(x * 7 + y * 31 + z * 101);
        if (h % 13u == 0u)
        {
          samples[static_cast<size_t>((x * ny + y) * nz + z)] = static_cast<double>(h % 5u);
          field.SetIndex(x, y, z, static_cast<double>(h % 5u));
        }
      }
  b200::ComputeDistanceFieldTransformInPlace(DegreeOfParallelism::None(), field);
  bool ok = true;
  for (int64_t x = 0; x < nx; x++)
    for (int64_t y = 0; y < ny; y++)
      for (int64_t z = 0; z < nz; z++)
      {
        double best = inf;
        for (int64_t a = 0; a < nx; a++)
          for (int64_t b = 0; b < ny; b++)
            for (int64_t c = 0; c < nz; c++)
            {
              const double candidate = samples[static_cast<size_t>((a * ny + b) * nz + c)]
                  + static_cast<double>((x - a) * (x - a) + (y - b) * (y - b) + (z - c) * (z - c));
              best = (candidate < best) ? candidate : best;
            }
        ok = ok && (field.GetIndexImmutable(x, y, z).Value() == best);
      }
  EXPECT_TRUE(ok);
  // a sample the exact integer passes cannot take -> runtime_error, field untouched
  field.SetIndex(0, 0, 0, 0.5);
  bool threw = false;
  try { b200::ComputeDistanceFieldTransformInPlace(DegreeOfParallelism::None(), field); }
  catch (const std::runtime_error&) { threw = true; }
  EXPECT_TRUE(threw && field.GetIndexImmutable(0, 0, 0).Value() == 0.5);
}

// The replacement factory (b200_voxelizer_backends.cpp) behind the reference's own declarations
// (pointcloud_voxelization.hpp:53-68), as test/pointcloud_voxelization_test.cpp:269-311 uses it:
// every available backend voxelizes the scene, and the B200 backend's map equals the map of the
// REFERENCE'S OWN CPU backend (cpu_pointcloud_voxelization.cpp, linked into this binary).
void FactoryTest()
{
  using namespace pointcloud_voxelization;
  const auto backends = GetAvailableBackends();
  EXPECT_TRUE(backends.size() >= 3);
  EXPECT_TRUE(backends.front().BackendOption() == BackendOptions::CUDA);
  EXPECT_TRUE(backends.front().DeviceOptions().at("CUDA_DEVICE") == 0);
  EXPECT_TRUE(backends.back().BackendOption() == BackendOptions::CPU);
  // a scene with clipped rays and rays from outside the grid
  const auto grid_sizes = VoxelGridSizes::FromGridSizes(0.125, Eigen::Vector3d(4.0, 3.0, 2.0));
  OccupancyMap static_environment(Eigen::Isometry3d::FromTranslation(-2.0, -1.5, -1.0), "world",
                                  grid_sizes, OccupancyCell(0.0f));
  for (int64_t x = 0; x < static_environment.NumXVoxels(); x++)
    for (int64_t y = 0; y < static_environment.NumYVoxels(); y++)
      static_environment.SetIndex(x, y, 0, OccupancyCell(1.0f));
  std::vector<PointCloudWrapperSharedPtr> clouds;
  const double poses[3][3] = {{-3.0, 0.1, 0.2}, {0.3, 0.2, 0.1}, {1.0, -2.5, 0.4}};
  for (int c = 0; c < 3; c++)
  {
    auto cloud = std::make_shared<VectorVector3dPointCloudWrapper>();
    cloud->SetPointCloudOriginTransform(
        Eigen::Isometry3d::FromTranslation(poses[c][0], poses[c][1], poses[c][2]));
    cloud->SetMaxRange((c == 1) ? 1.5 : std::numeric_limits<double>::infinity());
    uint32_t state = 12345u + static_cast<uint32_t>(c);
    for (int i = 0; i < 20000; i++)
    {
      double xyz[3];
      for (double& value : xyz)
      {
        state = state * 1664525u + 1013904223u;
        value = (static_cast<double>(state >> 8) / 16777216.0) * 6.0 - 3.0;
      }
      cloud->PushBack(xyz[0], xyz[1], xyz[2]);
    }
    clouds.push_back(cloud);
  }
  const PointCloudVoxelizationFilterOptions filter_options(0.9, 1, 1);
  std::vector<OccupancyMap> results;
  for (const auto& backend : backends)
  {
    std::vector<std::string> log;
    const auto voxelizer =
        MakePointCloudVoxelizer(backend, [&](const std::string& m) { log.push_back(m); });
    EXPECT_TRUE(voxelizer != nullptr);
    results.push_back(voxelizer->VoxelizePointClouds(static_environment, filter_options, clouds));
  }
  size_t filled = 0;
  for (const auto& cell : results.front().GetImmutableRawData())
  {
    filled += (cell.Occupancy() == 1.0f) ? 1u : 0u;
  }
  EXPECT_TRUE(filled > 500);
  for (size_t i = 1; i < results.size(); i++)
  {
    EXPECT_TRUE(std::memcmp(results.front().GetImmutableRawData().data(),
                            results[i].GetImmutableRawData().data(),
                            sizeof(float) * results[i].GetImmutableRawData().size()) == 0);
  }
  // BEST_AVAILABLE picks the device backend; OpenCL is not part of this build; null logging ok
  std::vector<std::string> log;
  const auto best = MakePointCloudVoxelizer(BackendOptions::BEST_AVAILABLE, {},
                                            [&](const std::string& m) { log.push_back(m); });
  EXPECT_TRUE(dynamic_cast<const B200PointCloudVoxelizer*>(best.get()) != nullptr);
  EXPECT_TRUE(!log.empty());
  bool threw = false;
  try { MakePointCloudVoxelizer(BackendOptions::OPENCL, {}); }
  catch (const std::runtime_error&) { threw = true; }
  EXPECT_TRUE(threw);
  EXPECT_TRUE(MakePointCloudVoxelizer(BackendOptions::CPU, {{"CPU_PARALLELIZE", 0}}) != nullptr);
}

// `adapter_test --time N`: the REAL drop-in call on an N^3 map - SignedDistanceField
// construction (the oob_value fill), the C-ABI call on pageable std::vector storage, Lock() -
// timed piece by piece; one JSON line for bench.py (e2e.cpp_adapter).
int TimeAdapter(const int64_t n)
{
  using Clock = std::chrono::steady_clock;
  const auto sizes = VoxelGridSizes::FromVoxelCounts(
      0.02, common_robotics_utilities::voxel_grid::Vector3i64(n, n, n));
  OccupancyMap map(Eigen::Isometry3d::Identity(), "bench", sizes, OccupancyCell(0.0f));
  {
    // blobs: ~10 % filled
    auto& cells = map.GetMutableRawData();
    const int64_t spheres = 96;
    uint32_t state = 42u;
    const auto next = [&]() { state = state * 1664525u + 1013904223u; return state >> 8; };
    for (int64_t s = 0; s < spheres; s++)
    {
      const int64_t cx = next() % n, cy = next() % n, cz = next() % n;
      const int64_t radius = n / 16 + static_cast<int64_t>(next() % (n / 12 + 1));
      for (int64_t x = std::max<int64_t>(0, cx - radius); x < std::min(n, cx + radius + 1); x++)
        for (int64_t y = std::max<int64_t>(0, cy - radius); y < std::min(n, cy + radius + 1); y++)
          for (int64_t z = std::max<int64_t>(0, cz - radius); z < std::min(n, cz + radius + 1); z++)
            if ((x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz) <= radius * radius)
              cells[static_cast<size_t>((x * n + y) * n + z)] = OccupancyCell(1.0f);
    }
  }
  const auto params = SDFGenerationParams<float>();
  const auto seconds = [](Clock::time_point a, Clock::time_point b)
  { return std::chrono::duration<double>(b - a).count(); };
  double whole = 0.0, construct = 0.0, call = 0.0, lock = 0.0, known = 0.0;
  const int reps = 3;
  for (int rep = 0; rep < reps + 1; rep++)
  {
    const auto t0 = Clock::now();
    const auto sdf = b200::ExtractSignedDistanceFieldFromOccupancyMap(map, params);
    const auto t1 = Clock::now();
    // the same pieces one by one
    SignedDistanceField<float> pieces(map.OriginTransform(), map.Frame(), map.ControlSizes(),
                                      params.OOBValue());
    const auto t2 = Clock::now();
    float lo = 0.0f, hi = 0.0f;
    b200::ThrowOnError(vgt_b200_sdf_f32(
        reinterpret_cast<const float*>(map.GetImmutableRawData().data()), n, n, n, 0.02, 1, 0, 0,
        pieces.GetMutableRawData().data(), &lo, &hi));
    const auto t3 = Clock::now();
    pieces.Lock();
    const auto t4 = Clock::now();
    pieces.Unlock();
#ifdef VGT_B200_SDF_HAS_LOCK_WITH_KNOWN_EXTREMA
    pieces.LockWithKnownExtrema(lo, hi);
#endif
    const auto t5 = Clock::now();
    if (!(pieces.GetImmutableRawData() == sdf.GetImmutableRawData())) { return 1; }
    if (rep > 0)
    {
      whole += seconds(t0, t1);
      construct += seconds(t1, t2);
      call += seconds(t2, t3);
      lock += seconds(t3, t4);
      known += seconds(t4, t5);
    }
  }
  const double voxels = static_cast<double>(n) * n * n;
  std::printf(
      "{\"api\": \"b200::ExtractSignedDistanceFieldFromOccupancyMap (std::vector storage, "
      "SignedDistanceField construction, C-ABI call, Lock)\", \"grid\": \"%lld^3\", "
      "\"ms_per_call\": %.3f, \"gvoxels_per_s\": %.3f, \"construct_fill_ms\": %.3f, "
      "\"c_abi_call_pageable_ms\": %.3f, \"lock_minmax_scan_ms\": %.3f, "
      "\"lock_with_known_extrema_ms\": %.4f, \"fast_lock_compiled_in\": %s}\n",
      static_cast<long long>(n), whole / reps * 1e3, voxels / (whole / reps) / 1e9,
      construct / reps * 1e3, call / reps * 1e3, lock / reps * 1e3, known / reps * 1e3,
#ifdef VGT_B200_SDF_HAS_LOCK_WITH_KNOWN_EXTREMA
      "true"
#else
      "false"
#endif
  );
  return 0;
}

// Port of test/mesh_rasterization_test.cpp (TestOccupancyMap :20-66, TestOccupancyComponentMap
// :68-114) through the adapter, then the adapter against the reference's own CPU rasterizer
// (linked in from the reference's mesh_rasterizer.cpp) on a closed mesh, cell for cell.
template <typename MapType>
void CheckReferenceTriangle(const MapType& occupancy_map)
{
  const auto occupancy = [&](const int64_t x, const int64_t y, const int64_t z)
  { return occupancy_map.GetIndexImmutable(x, y, z).Value().Occupancy(); };
  for (int64_t x = 0; x < occupancy_map.NumXVoxels(); x++)
  {
    for (int64_t y = 0; y < occupancy_map.NumYVoxels(); y++)
    {
      EXPECT_TRUE(occupancy(x, y, 0) == 0.0f);
      if (x == 0 || y == 0 || y >= (occupancy_map.NumYVoxels() - x))
      {
        EXPECT_TRUE(occupancy(x, y, 1) == 0.0f);
      }
      else
      {
        EXPECT_TRUE(occupancy(x, y, 1) == 1.0f);
      }
    }
  }
}

void MeshRasterizerTest()
{
  namespace gpu = mesh_rasterizer::b200;
  const DegreeOfParallelism parallelism = DegreeOfParallelism::None();
  const std::vector<Eigen::Vector3d> vertices = {
      Eigen::Vector3d(0.0, 0.0, 0.0), Eigen::Vector3d(1.0, 0.0, 0.0),
      Eigen::Vector3d(0.0, 1.0, 0.0)};
  const std::vector<Eigen::Vector3i> triangles = {Eigen::Vector3i(0, 1, 2)};
  CheckReferenceTriangle(
      gpu::RasterizeMeshIntoOccupancyMap(vertices, triangles, 0.125, parallelism));
  CheckReferenceTriangle(
      gpu::RasterizeMeshIntoOccupancyComponentMap(vertices, triangles, 0.125, parallelism));

  // an octahedron with its faces split once: device == the reference's CPU rasterizer
  std::vector<Eigen::Vector3d> solid = {
      Eigen::Vector3d(0.7, 0.0, 0.0), Eigen::Vector3d(-0.6, 0.0, 0.1),
      Eigen::Vector3d(0.0, 0.5, 0.0), Eigen::Vector3d(0.1, -0.8, 0.0),
      Eigen::Vector3d(0.0, 0.05, 0.9), Eigen::Vector3d(0.0, 0.0, -0.4)};
  const std::vector<Eigen::Vector3i> faces = {
      Eigen::Vector3i(0, 2, 4), Eigen::Vector3i(2, 1, 4), Eigen::Vector3i(1, 3, 4),
      Eigen::Vector3i(3, 0, 4), Eigen::Vector3i(2, 0, 5), Eigen::Vector3i(1, 2, 5),
      Eigen::Vector3i(3, 1, 5), Eigen::Vector3i(0, 3, 5)};
  for (const double resolution : {0.05, 0.03125})
  {
    const auto device_map = gpu::RasterizeMeshIntoOccupancyMap(solid, faces, resolution,
                                                               parallelism);
    const auto cpu_map = mesh_rasterizer::RasterizeMeshIntoOccupancyMap(solid, faces, resolution,
                                                                        parallelism);
    EXPECT_TRUE(device_map.NumXVoxels() == cpu_map.NumXVoxels()
                && device_map.NumYVoxels() == cpu_map.NumYVoxels()
                && device_map.NumZVoxels() == cpu_map.NumZVoxels());
    int64_t filled = 0, different = 0;
    for (int64_t i = 0; i < cpu_map.NumTotalVoxels(); i++)
    {
      const float want = cpu_map.GetDataIndexImmutable(i).Occupancy();
      filled += (want == 1.0f) ? 1 : 0;
      different += (device_map.GetDataIndexImmutable(i).Occupancy() != want) ? 1 : 0;
    }
    EXPECT_TRUE(filled > 500);
    EXPECT_TRUE(different == 0);
  }

  // RasterizeMesh / RasterizeTriangle into an existing map; errors as the reference throws them
  const auto sizes = VoxelGridSizes::FromGridSizes(0.1, Eigen::Vector3d(1.0, 1.0, 1.0));
  OccupancyMap small_map(Eigen::Isometry3d::FromTranslation(0.0, 0.0, 0.0), "world", sizes,
                         OccupancyCell(0.0f));
  OccupancyMap cpu_small = small_map;
  gpu::RasterizeMesh(solid, faces, small_map, false, parallelism);
  mesh_rasterizer::RasterizeMesh(solid, faces, cpu_small, false, parallelism);
  int64_t different = 0;
  for (int64_t i = 0; i < cpu_small.NumTotalVoxels(); i++)
  {
    different += (small_map.GetDataIndexImmutable(i).Occupancy()
                  != cpu_small.GetDataIndexImmutable(i).Occupancy()) ? 1 : 0;
  }
  EXPECT_TRUE(different == 0);
  bool threw = false;
  try { gpu::RasterizeMesh(solid, faces, small_map, true, parallelism); }
  catch (const std::runtime_error&) { threw = true; }
  EXPECT_TRUE(threw);        // the solid leaves the unit map
  threw = false;
  try { gpu::RasterizeTriangle(solid, faces, faces.size(), small_map, false); }
  catch (const std::out_of_range&) { threw = true; }
  EXPECT_TRUE(threw);
  threw = false;
  try { gpu::RasterizeMesh(solid, {Eigen::Vector3i(0, 1, 6)}, small_map, false, parallelism); }
  catch (const std::out_of_range&) { threw = true; }
  EXPECT_TRUE(threw);
  threw = false;
  try { gpu::RasterizeMeshIntoOccupancyMap(solid, faces, 0.0, parallelism); }
  catch (const std::invalid_argument&) { threw = true; }
  EXPECT_TRUE(threw);
}

// the SDF files. The reference's own SaveToFile / LoadFromFile (compiled here over
// the stand-in serialization layer) against the library's writer / reader: same bytes, and
// each side loads what the other wrote; a field computed on the device goes to a file
// without a host copy of the caller's own.
std::vector<char> FileBytes(const std::string& path)
{
  std::vector<char> bytes;
  if (FILE* file = std::fopen(path.c_str(), "rb"))
  {
    char buffer[4096];
    size_t got = 0;
    while ((got = std::fread(buffer, 1, sizeof(buffer), file)) > 0)
    {
      bytes.insert(bytes.end(), buffer, buffer + got);
    }
    std::fclose(file);
  }
  return bytes;
}

template <typename ScalarType>
void GridFileTest()
{
  namespace files = grid_files::b200;
  OccupancyMap map = MakeMap(0.25, 2.0, 1.5, 1.0, 0.0f);
  for (int64_t x = 2; x < 5; x++)
  {
    for (int64_t y = 1; y < 4; y++)
    {
      map.SetIndex(x, y, 2, OccupancyCell(1.0f));
    }
  }
  const SignedDistanceField<ScalarType> sdf =
      b200::ExtractSignedDistanceFieldFromOccupancyMap<OccupancyMap, ScalarType>(
          map, SDFGenerationParams<ScalarType>());
  const std::string base = std::string("/tmp/vgt_b200_adapter_") + (sizeof(ScalarType) == 4 ? "f" : "d");
  for (const bool compress : {false, true})
  {
    const std::string theirs = base + (compress ? "_theirs.sdfz" : "_theirs.sdfr");
    const std::string ours = base + (compress ? "_ours.sdfz" : "_ours.sdfr");
    SignedDistanceField<ScalarType>::SaveToFile(sdf, theirs, compress);
    files::SaveSignedDistanceFieldToFile(sdf, ours, compress);
    const auto their_bytes = FileBytes(theirs);
    EXPECT_TRUE(!their_bytes.empty());
    EXPECT_TRUE(their_bytes == FileBytes(ours));
    // the reference's reader on our file, our reader on the reference's file
    const auto from_ours = SignedDistanceField<ScalarType>::LoadFromFile(ours);
    const auto from_theirs = files::LoadSignedDistanceFieldFromFile<ScalarType>(theirs);
    for (const auto* loaded : {&from_ours, &from_theirs})
    {
      EXPECT_TRUE(loaded->GetImmutableRawData() == sdf.GetImmutableRawData());
      EXPECT_TRUE(loaded->Frame() == sdf.Frame());
      EXPECT_TRUE(loaded->IsLocked() == sdf.IsLocked());
      EXPECT_TRUE(loaded->NumXVoxels() == sdf.NumXVoxels() && loaded->NumZVoxels() == sdf.NumZVoxels());
      EXPECT_TRUE(loaded->GetMinimumMaximum().Minimum() == sdf.GetMinimumMaximum().Minimum());
      EXPECT_TRUE(loaded->GetMinimumMaximum().Maximum() == sdf.GetMinimumMaximum().Maximum());
    }
    std::remove(theirs.c_str());
    std::remove(ours.c_str());
  }
  bool threw = false;
  try
  {
    files::LoadSignedDistanceFieldFromFile<ScalarType>(base + "_missing.sdf");
  }
  catch (const std::invalid_argument& error)
  {
    threw = std::string(error.what()) == "File does not exist";
  }
  EXPECT_TRUE(threw);
}

int main(int argc, char** argv)
{
  if (vgt_b200_device_count() < 1)
  {
    std::printf("no usable CUDA device: %s\n", vgt_b200_version());
    return 2;
  }
  if (argc >= 3 && std::string(argv[1]) == "--time")
  {
    return TimeAdapter(std::atoll(argv[2]));
  }
  SdfTests();
  MultiDeviceTest();
  VoxelizationTest();
  CellMapTests<float>();
  CellMapTests<double>();
  TransformTest();
  FactoryTest();
  MeshRasterizerTest();
  GridFileTest<float>();
  GridFileTest<double>();
  if (g_failures == 0)
  {
    std::printf("ADAPTER_TEST_OK\n");
    return 0;
  }
  std::printf("%d expectation(s) failed\n", g_failures);
  return 1;
}
