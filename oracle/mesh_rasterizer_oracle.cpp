// =============================================================================
// TEST INFRASTRUCTURE ONLY -- CPU oracle for the mesh rasterizer.
//
// From-scratch restatement of the reference's triangle rasterizer. Only tests/,
// bench.py's cpu_baseline leg and __graft_entry__.smoke() may use it; the
// product path never does.
//
// Parity status: PINNED. tests/test_mesh_rasterizer_vs_reference.py checks it
// bit for bit against the reference's own mesh_rasterizer.cpp compiled
// unmodified into oracle/_ref (over the stand-in Eigen / common_robotics_utilities
// headers of oracle/ref_shim, which fix the operation order below), and against
// the expectations of the reference's own test (test/mesh_rasterization_test.cpp:20-66).
//
// What it follows (paths relative to the reference checkout):
//   src/voxelized_geometry_tools/mesh_rasterizer.cpp:26-44    PointProjectsInsideTriangle
//   src/voxelized_geometry_tools/mesh_rasterizer.cpp:46-58    ClosestPointOnLineSegment
//   src/voxelized_geometry_tools/mesh_rasterizer.cpp:60-103   CalcClosestPointOnTriangle
//       (note: the edge case picks among the three edge points by THEIR OWN
//        squared norms, not by their distance to the query -- mirrored as written)
//   src/voxelized_geometry_tools/mesh_rasterizer.cpp:105-203  RasterizeTriangleImpl
//   src/voxelized_geometry_tools/mesh_rasterizer.cpp:205-230  RasterizeMeshImpl
//   src/voxelized_geometry_tools/mesh_rasterizer.cpp:232-279  RasterizeMeshInto...MapImpl
//
// Third-party arithmetic restated (Eigen and common_robotics_utilities are not
// vendored in the reference): dot = (ax*bx + ay*by) + az*bz; cross by components;
// VectorRejection(n, v) = v - n * (n.v / n.n) (zero projection for n.n == 0);
// ClampValue = min(hi, max(lo, v)); location -> index floor(p * (1 / voxel));
// index -> centre voxel * (i + 0.5); X * p per row ((m0*x + m1*y) + m2*z) + m3*w.
// No FMA contraction: build with -ffp-contract=off.
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>

namespace
{
struct V3
{
  double x, y, z;
};

inline V3 Sub(const V3& a, const V3& b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 Add(const V3& a, const V3& b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 Scale(const V3& a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline double Dot(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 Cross(const V3& a, const V3& b)
{
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// mesh_rasterizer.cpp:30-38
inline bool SameSide(const V3& a, const V3& b, const V3& p1, const V3& p2)
{
  const V3 ab = Sub(b, a);
  const V3 cross1 = Cross(ab, Sub(p1, a));
  const V3 cross2 = Cross(ab, Sub(p2, a));
  return Dot(cross1, cross2) >= 0.0;
}

// mesh_rasterizer.cpp:46-58
inline V3 ClosestOnSegment(const V3& a, const V3& b, const V3& q)
{
  const V3 ab = Sub(b, a);
  const V3 aq = Sub(q, a);
  const double ratio = Dot(ab, aq) / Dot(ab, ab);
  const double clamped = std::min(1.0, std::max(0.0, ratio));
  return Add(a, Scale(ab, clamped));
}

// mesh_rasterizer.cpp:60-103
inline V3 ClosestOnTriangle(const V3& v1, const V3& v2, const V3& v3, const V3& normal,
                            const V3& q)
{
  if (SameSide(v1, v2, v3, q) && SameSide(v2, v3, v1, q) && SameSide(v3, v1, v2, q))
  {
    const V3 v1q = Sub(q, v1);
    const double normal_squared = Dot(normal, normal);
    V3 projection{0.0, 0.0, 0.0};
    if (normal_squared > 0.0)
    {
      projection = Scale(normal, Dot(normal, v1q) / normal_squared);
    }
    return Add(v1, Sub(v1q, projection));
  }
  const V3 c12 = ClosestOnSegment(v1, v2, q);
  const V3 c23 = ClosestOnSegment(v2, v3, q);
  const V3 c31 = ClosestOnSegment(v3, v1, q);
  const double d12 = Dot(c12, c12);
  const double d23 = Dot(c23, c23);
  const double d31 = Dot(c31, c31);
  if (d12 <= d23 && d12 <= d31) { return c12; }
  if (d23 <= d12 && d23 <= d31) { return c23; }
  return c31;
}

inline V3 Transform(const double* m, const V3& p)   // column-major 4x4, w = 1
{
  return V3{((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12] * 1.0,
            ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13] * 1.0,
            ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14] * 1.0};
}
}  // namespace

extern "C"
{
// RasterizeMesh into `occupancy` (float [nx*ny*nz], x slowest; intersected cells set to 1.0).
// x_wg = the map's origin transform, x_gw = its inverse, both column-major 4x4.
// Returns 0; 2 when enforce_contains is set and an intersected cell lies outside the map
// (the reference throws std::runtime_error at the first such cell; here every cell is still
// processed); 3 when a triangle names a vertex that does not exist (std::out_of_range).
int vgt_oracle_rasterize_mesh_f64(const double* vertices, int64_t num_vertices,
                                  const int32_t* triangles, int64_t num_triangles,
                                  float* occupancy, int64_t nx, int64_t ny, int64_t nz,
                                  double resolution, const double* x_wg, const double* x_gw,
                                  int enforce_contains)
{
  const double min_check_radius = resolution * 0.5;
  const double max_check_radius = min_check_radius * std::sqrt(3.0);
  const double max_check_radius_squared = std::pow(max_check_radius, 2.0);
  const double inverse = 1.0 / resolution;
  int code = 0;
  for (int64_t t = 0; t < num_triangles; t++)
  {
    V3 v[3];
    for (int k = 0; k < 3; k++)
    {
      const int64_t index = triangles[3 * t + k];
      if (index < 0 || index >= num_vertices) { return 3; }
      v[k] = V3{vertices[3 * index], vertices[3 * index + 1], vertices[3 * index + 2]};
    }
    const V3 normal = Cross(Sub(v[1], v[0]), Sub(v[2], v[0]));
    const V3 low{std::min({v[0].x, v[1].x, v[2].x}), std::min({v[0].y, v[1].y, v[2].y}),
                 std::min({v[0].z, v[1].z, v[2].z})};
    const V3 high{std::max({v[0].x, v[1].x, v[2].x}), std::max({v[0].y, v[1].y, v[2].y}),
                  std::max({v[0].z, v[1].z, v[2].z})};
    const V3 low_grid = Transform(x_gw, low);
    const V3 high_grid = Transform(x_gw, high);
    const int64_t lo[3] = {static_cast<int64_t>(std::floor(low_grid.x * inverse)),
                           static_cast<int64_t>(std::floor(low_grid.y * inverse)),
                           static_cast<int64_t>(std::floor(low_grid.z * inverse))};
    const int64_t hi[3] = {static_cast<int64_t>(std::floor(high_grid.x * inverse)),
                           static_cast<int64_t>(std::floor(high_grid.y * inverse)),
                           static_cast<int64_t>(std::floor(high_grid.z * inverse))};
    for (int64_t x = lo[0]; x <= hi[0]; x++)
    {
      for (int64_t y = lo[1]; y <= hi[1]; y++)
      {
        for (int64_t z = lo[2]; z <= hi[2]; z++)
        {
          const V3 centre{resolution * (static_cast<double>(x) + 0.5),
                          resolution * (static_cast<double>(y) + 0.5),
                          resolution * (static_cast<double>(z) + 0.5)};
          const V3 q = Transform(x_wg, centre);
          const V3 closest = ClosestOnTriangle(v[0], v[1], v[2], normal, q);
          const V3 offset = Sub(closest, q);
          if (Dot(offset, offset) <= max_check_radius_squared)
          {
            if (x >= 0 && x < nx && y >= 0 && y < ny && z >= 0 && z < nz)
            {
              occupancy[(x * ny + y) * nz + z] = 1.0f;
            }
            else if (enforce_contains)
            {
              code = 2;
            }
          }
        }
      }
    }
  }
  return code;
}

// The map RasterizeMeshIntoOccupancyMap sizes around a mesh (mesh_rasterizer.cpp:243-272):
// dims[3] and the origin translation[3] (identity rotation). Returns 4 for resolution <= 0.
int vgt_oracle_mesh_map_extent_f64(const double* vertices, int64_t num_vertices, double resolution,
                                   int64_t* dims, double* origin_translation)
{
  if (resolution <= 0.0) { return 4; }
  const double infinity = std::numeric_limits<double>::infinity();
  double low[3] = {infinity, infinity, infinity};
  double high[3] = {-infinity, -infinity, -infinity};
  for (int64_t i = 0; i < num_vertices; i++)
  {
    for (int k = 0; k < 3; k++)
    {
      const double value = vertices[3 * i + k];
      low[k] = low[k] < value ? low[k] : value;       // cwiseMin(self, vertex)
      high[k] = high[k] > value ? high[k] : value;
    }
  }
  const double buffer_size = resolution * 2.0;
  for (int k = 0; k < 3; k++)
  {
    const double extent = (high[k] - low[k]) + buffer_size;
    dims[k] = static_cast<int64_t>(std::ceil(extent / resolution));
    origin_translation[k] = low[k] - resolution;
  }
  return 0;
}
}  // extern "C"
