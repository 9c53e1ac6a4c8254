// TEST INFRASTRUCTURE ONLY -- C entry points over the reference's own mesh rasterizer
// (src/voxelized_geometry_tools/mesh_rasterizer.cpp, compiled unmodified over the stand-in
// third-party headers of oracle/ref_shim). Used by tests/test_mesh_rasterizer_vs_reference.py to
// pin oracle/mesh_rasterizer_oracle.cpp, and by the C++ adapter test as the CPU side.
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include <voxelized_geometry_tools/mesh_rasterizer.hpp>

namespace
{
namespace vgt = voxelized_geometry_tools;
using common_robotics_utilities::parallelism::DegreeOfParallelism;
using common_robotics_utilities::voxel_grid::Vector3i64;
using common_robotics_utilities::voxel_grid::VoxelGridSizes;

void ReadMesh(const double* vertices, int64_t num_vertices, const int32_t* triangles,
              int64_t num_triangles, std::vector<Eigen::Vector3d>& vertex_list,
              std::vector<Eigen::Vector3i>& triangle_list)
{
  for (int64_t i = 0; i < num_vertices; i++)
  {
    vertex_list.emplace_back(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
  }
  for (int64_t i = 0; i < num_triangles; i++)
  {
    triangle_list.emplace_back(triangles[3 * i], triangles[3 * i + 1], triangles[3 * i + 2]);
  }
}

DegreeOfParallelism Threads(int threads)
{
  if (threads <= 0) { return DegreeOfParallelism::FromOmp(); }
  return DegreeOfParallelism(threads);
}
}  // namespace

extern "C"
{
// RasterizeMesh into an existing map (occupancy in/out). 0 ok, 2 = the reference threw
// std::runtime_error (a triangle leaves the map while enforce is set), 3 = std::out_of_range
// (a triangle names a vertex that does not exist), 1 anything else.
int vgt_ref_rasterize_mesh(const double* vertices, int64_t num_vertices, const int32_t* triangles,
                           int64_t num_triangles, float* occupancy, int64_t nx, int64_t ny,
                           int64_t nz, double resolution, const double* origin_column_major,
                           int enforce_contains, int threads)
{
  try
  {
    std::vector<Eigen::Vector3d> vertex_list;
    std::vector<Eigen::Vector3i> triangle_list;
    ReadMesh(vertices, num_vertices, triangles, num_triangles, vertex_list, triangle_list);
    Eigen::Isometry3d origin;
    std::memcpy(origin.data(), origin_column_major, sizeof(double) * 16);
    const auto sizes = VoxelGridSizes::FromVoxelCounts(resolution, Vector3i64(nx, ny, nz));
    vgt::OccupancyMap map(origin, "mesh", sizes, vgt::OccupancyCell(0.0f));
    static_assert(sizeof(vgt::OccupancyCell) == sizeof(float), "cell layout");
    std::memcpy(static_cast<void*>(map.GetMutableRawData().data()), occupancy,
                sizeof(float) * static_cast<size_t>(nx * ny * nz));
    int code = 0;
    try
    {
      vgt::mesh_rasterizer::RasterizeMesh(vertex_list, triangle_list, map, enforce_contains != 0,
                                          Threads(threads));
    }
    catch (const std::out_of_range&) { code = 3; }
    catch (const std::runtime_error&) { code = 2; }
    std::memcpy(occupancy, static_cast<const void*>(map.GetImmutableRawData().data()),
                sizeof(float) * static_cast<size_t>(nx * ny * nz));
    return code;
  }
  catch (const std::exception&)
  {
    return 1;
  }
}

// RasterizeMeshIntoOccupancyMap: the map it sizes around the mesh. dims / origin always written;
// the occupancy only when `capacity` voxels are enough.
int vgt_ref_rasterize_mesh_into_map(const double* vertices, int64_t num_vertices,
                                    const int32_t* triangles, int64_t num_triangles,
                                    double resolution, int threads, int64_t* dims,
                                    double* origin_column_major, float* occupancy,
                                    int64_t capacity)
{
  try
  {
    std::vector<Eigen::Vector3d> vertex_list;
    std::vector<Eigen::Vector3i> triangle_list;
    ReadMesh(vertices, num_vertices, triangles, num_triangles, vertex_list, triangle_list);
    const auto map = vgt::mesh_rasterizer::RasterizeMeshIntoOccupancyMap(
        vertex_list, triangle_list, resolution, Threads(threads));
    dims[0] = map.NumXVoxels();
    dims[1] = map.NumYVoxels();
    dims[2] = map.NumZVoxels();
    std::memcpy(origin_column_major, map.OriginTransform().data(), sizeof(double) * 16);
    if (capacity >= map.NumTotalVoxels())
    {
      std::memcpy(occupancy, static_cast<const void*>(map.GetImmutableRawData().data()),
                  sizeof(float) * static_cast<size_t>(map.NumTotalVoxels()));
    }
    return 0;
  }
  catch (const std::invalid_argument&) { return 4; }
  catch (const std::runtime_error&) { return 2; }
  catch (const std::exception&) { return 1; }
}
}  // extern "C"
