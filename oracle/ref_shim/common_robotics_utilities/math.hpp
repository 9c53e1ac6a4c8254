// TEST INFRASTRUCTURE ONLY -- stand-in for common_robotics_utilities/math.hpp with the one
// function signed_distance_field.hpp calls and the one mesh_rasterizer.cpp calls. The real library is not in the reference tree
// (unvendored, unpinned), so this blend is a RESTATEMENT, the same one that
// oracle/sdf_queries_oracle.py and csrc/sdf_queries.cu state: per axis the ratio
// (query - low) / (high - low) clamped to [0, 1] (NaN passes through), each blend in the form
// a * (1 - ratio) + b * ratio, along x, then y, then z. It is the part of the SDF queries that
// stays parity-unpinned (tolerance 1e-12 relative, oracle/sdf_queries_oracle.py).
#pragma once

#include <vector>

#include <Eigen/Geometry>

namespace common_robotics_utilities
{
namespace math
{
// (test/pointcloud_voxelization_test.cpp:78: a cloud as a vector of points)
using VectorVector3d = std::vector<Eigen::Vector3d>;

inline double ClampedRatio(const double query, const double low, const double high)
{
  const double ratio = (query - low) / (high - low);
  if (ratio != ratio) { return ratio; }
  return ratio < 0.0 ? 0.0 : (ratio > 1.0 ? 1.0 : ratio);
}

template <typename T>
inline T Interpolate(const T& p1, const T& p2, const double ratio)
{
  return (p1 * (1.0 - ratio)) + (p2 * ratio);
}
// (scalars by value, so that a constexpr local can be passed from a capture-less lambda:
// test/voxel_raycasting_test.cpp:40-45)
inline double Interpolate(const double p1, const double p2, const double ratio)
{
  return (p1 * (1.0 - ratio)) + (p2 * ratio);
}

template <typename T>
inline T TrilinearInterpolate(
    const Eigen::Vector3d& low_corner, const Eigen::Vector3d& high_corner,
    const T& mxmymz, const T& mxmypz, const T& mxpymz, const T& mxpypz,
    const T& pxmymz, const T& pxmypz, const T& pxpymz, const T& pxpypz,
    const Eigen::Vector3d& query)
{
  const double rx = ClampedRatio(query(0), low_corner(0), high_corner(0));
  const double ry = ClampedRatio(query(1), low_corner(1), high_corner(1));
  const double rz = ClampedRatio(query(2), low_corner(2), high_corner(2));
  const T mm = Interpolate(mxmymz, pxmymz, rx);
  const T mp = Interpolate(mxmypz, pxmypz, rx);
  const T pm = Interpolate(mxpymz, pxpymz, rx);
  const T pp = Interpolate(mxpypz, pxpypz, rx);
  const T m = Interpolate(mm, pm, ry);
  const T p = Interpolate(mp, pp, ry);
  return Interpolate(m, p, rz);
}
// VectorProjection / VectorRejection (restated like the blend above: the component of `vector`
// along `base_vector` is (base . vector / base . base) * base, zero for a zero base; the
// rejection is vector minus that).
inline Eigen::Vector3d VectorProjection(const Eigen::Vector3d& base_vector,
                                        const Eigen::Vector3d& vector)
{
  const double base_squared_norm = base_vector.squaredNorm();
  if (base_squared_norm > 0.0)
  {
    return base_vector * (base_vector.dot(vector) / base_squared_norm);
  }
  return Eigen::Vector3d(0.0, 0.0, 0.0);
}

inline Eigen::Vector3d VectorRejection(const Eigen::Vector3d& base_vector,
                                       const Eigen::Vector3d& vector)
{
  return vector - VectorProjection(base_vector, vector);
}
}  // namespace math
}  // namespace common_robotics_utilities
