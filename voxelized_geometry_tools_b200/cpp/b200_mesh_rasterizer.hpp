// C++ host adapter: the reference's mesh rasterizer signatures
// (include/voxelized_geometry_tools/mesh_rasterizer.hpp:19-91), bodies marshalled to the B200
// C-ABI (include/vgt_b200.h: vgt_b200_rasterize_mesh_f64). Header-only; compiles against the
// reference's own headers. Same names, argument meaning and error behaviour, in namespace
// voxelized_geometry_tools::mesh_rasterizer::b200 so it can sit next to the CPU implementation
// while a maintainer switches call sites.
#pragma once

#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include <Eigen/Geometry>
#include <common_robotics_utilities/parallelism.hpp>
#include <common_robotics_utilities/voxel_grid.hpp>
#include <voxelized_geometry_tools/occupancy_component_map.hpp>
#include <voxelized_geometry_tools/occupancy_map.hpp>
#include <voxelized_geometry_tools/vgt_namespace.hpp>

#include "vgt_b200.h"

namespace voxelized_geometry_tools
{
VGT_NAMESPACE_BEGIN
namespace mesh_rasterizer
{
namespace b200
{
namespace internal
{
// C-ABI status -> the exception the reference throws at the same place
// (mesh_rasterizer.cpp:113-116 invalid_argument, :122-125 out_of_range, :191-196 runtime_error).
inline void ThrowOnError(const int status)
{
  if (status == VGT_B200_OK)
  {
    return;
  }
  const std::string message = vgt_b200_last_error();
  if (status == VGT_B200_ERR_INVALID_ARGUMENT)
  {
    throw std::invalid_argument(message);
  }
  if (status == VGT_B200_ERR_OUT_OF_RANGE)
  {
    throw std::out_of_range(message);
  }
  throw std::runtime_error(message);
}

template <typename OccupancyMapType>
inline void RasterizeTriangles(
    const std::vector<Eigen::Vector3d>& vertices, const Eigen::Vector3i* triangles,
    const size_t num_triangles, OccupancyMapType& occupancy_map, const bool enforce_contains,
    const int device)
{
  if (!occupancy_map.IsInitialized())
  {
    throw std::invalid_argument("occupancy_map must be initialized");
  }
  std::vector<double> vertices_xyz(vertices.size() * 3);
  for (size_t i = 0; i < vertices.size(); i++)
  {
    vertices_xyz[3 * i] = vertices[i].x();
    vertices_xyz[3 * i + 1] = vertices[i].y();
    vertices_xyz[3 * i + 2] = vertices[i].z();
  }
  std::vector<int32_t> triangle_indices(num_triangles * 3);
  for (size_t i = 0; i < num_triangles; i++)
  {
    triangle_indices[3 * i] = triangles[i](0);
    triangle_indices[3 * i + 1] = triangles[i](1);
    triangle_indices[3 * i + 2] = triangles[i](2);
  }
  using CellType = typename std::remove_reference<
      decltype(occupancy_map.GetMutableRawData())>::type::value_type;
  static_assert(sizeof(CellType) == 4 || sizeof(CellType) == 8,
                "OccupancyCell (4 bytes) or OccupancyComponentCell (8 bytes) expected");
  const Eigen::Isometry3d origin_transform = occupancy_map.OriginTransform();
  const Eigen::Isometry3d inverse_origin_transform = occupancy_map.InverseOriginTransform();
  ThrowOnError(vgt_b200_rasterize_mesh_f64(
      vertices_xyz.data(), static_cast<int64_t>(vertices.size()), triangle_indices.data(),
      static_cast<int64_t>(num_triangles),
      static_cast<void*>(occupancy_map.GetMutableRawData().data()),
      static_cast<int>(sizeof(CellType)), occupancy_map.NumXVoxels(), occupancy_map.NumYVoxels(),
      occupancy_map.NumZVoxels(), occupancy_map.Resolution(), origin_transform.data(),
      inverse_origin_transform.data(), enforce_contains ? 1 : 0, device));
}

template <typename OccupancyMapType, typename OccupancyCellType>
inline OccupancyMapType RasterizeMeshIntoMap(
    const std::vector<Eigen::Vector3d>& vertices, const std::vector<Eigen::Vector3i>& triangles,
    const double resolution, const int device)
{
  // mesh_rasterizer.cpp:239-272: the mesh's bounding box plus one voxel on every side
  if (resolution <= 0.0)
  {
    throw std::invalid_argument("resolution must be greater than zero");
  }
  Eigen::Vector3d lower_corner =
      Eigen::Vector3d::Constant(std::numeric_limits<double>::infinity());
  Eigen::Vector3d upper_corner =
      Eigen::Vector3d::Constant(-std::numeric_limits<double>::infinity());
  for (const Eigen::Vector3d& vertex : vertices)
  {
    lower_corner = lower_corner.cwiseMin(vertex);
    upper_corner = upper_corner.cwiseMax(vertex);
  }
  const Eigen::Vector3d object_size = upper_corner - lower_corner;
  const double buffer_size = resolution * 2.0;
  const Eigen::Vector3d grid_dimensions(object_size.x() + buffer_size,
                                        object_size.y() + buffer_size,
                                        object_size.z() + buffer_size);
  const auto grid_sizes = common_robotics_utilities::voxel_grid::VoxelGridSizes::FromGridSizes(
      resolution, grid_dimensions);
  const Eigen::Isometry3d origin_transform(Eigen::Translation3d(
      lower_corner.x() - resolution, lower_corner.y() - resolution,
      lower_corner.z() - resolution));
  OccupancyMapType occupancy_map(origin_transform, "mesh", grid_sizes, OccupancyCellType(0.0f));
  RasterizeTriangles(vertices, triangles.data(), triangles.size(), occupancy_map, true, device);
  return occupancy_map;
}
}  // namespace internal

// mesh_rasterizer.hpp:27-32 / :42-47
template <typename OccupancyMapType>
inline void RasterizeTriangle(
    const std::vector<Eigen::Vector3d>& vertices, const std::vector<Eigen::Vector3i>& triangles,
    const size_t triangle_index, OccupancyMapType& occupancy_map,
    const bool enforce_occupancy_map_contains_triangle, const int device = 0)
{
  internal::RasterizeTriangles(vertices, &triangles.at(triangle_index), 1, occupancy_map,
                               enforce_occupancy_map_contains_triangle, device);
}

// mesh_rasterizer.hpp:52-58 / :63-69 (`parallelism` is accepted and ignored: the device decides)
template <typename OccupancyMapType>
inline void RasterizeMesh(
    const std::vector<Eigen::Vector3d>& vertices, const std::vector<Eigen::Vector3i>& triangles,
    OccupancyMapType& occupancy_map, const bool enforce_occupancy_map_contains_mesh,
    const common_robotics_utilities::parallelism::DegreeOfParallelism& /* parallelism */,
    const int device = 0)
{
  internal::RasterizeTriangles(vertices, triangles.data(), triangles.size(), occupancy_map,
                               enforce_occupancy_map_contains_mesh, device);
}

// mesh_rasterizer.hpp:75-79
inline OccupancyMap RasterizeMeshIntoOccupancyMap(
    const std::vector<Eigen::Vector3d>& vertices, const std::vector<Eigen::Vector3i>& triangles,
    const double resolution,
    const common_robotics_utilities::parallelism::DegreeOfParallelism& /* parallelism */,
    const int device = 0)
{
  return internal::RasterizeMeshIntoMap<OccupancyMap, OccupancyCell>(vertices, triangles,
                                                                     resolution, device);
}

// mesh_rasterizer.hpp:85-89
inline OccupancyComponentMap RasterizeMeshIntoOccupancyComponentMap(
    const std::vector<Eigen::Vector3d>& vertices, const std::vector<Eigen::Vector3i>& triangles,
    const double resolution,
    const common_robotics_utilities::parallelism::DegreeOfParallelism& /* parallelism */,
    const int device = 0)
{
  return internal::RasterizeMeshIntoMap<OccupancyComponentMap, OccupancyComponentCell>(
      vertices, triangles, resolution, device);
}
}  // namespace b200
}  // namespace mesh_rasterizer
VGT_NAMESPACE_END
}  // namespace voxelized_geometry_tools
