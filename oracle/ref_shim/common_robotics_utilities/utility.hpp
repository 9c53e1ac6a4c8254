// TEST INFRASTRUCTURE ONLY -- stand-in for common_robotics_utilities/utility.hpp with the one
// type the reference's CPU voxelizer uses: an atomic that can be copied (a std::vector of
// tracking grids is built from one prototype, cpu_pointcloud_voxelization.cpp:146-149).
#pragma once

#include <algorithm>
#include <atomic>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

namespace common_robotics_utilities
{
namespace utility
{
// (mesh_rasterizer.cpp:55: restated as min(max, max(min, value)), so a NaN value gives `min`)
template <typename T>
inline T ClampValue(const T& value, const T& min, const T& max)
{
  if (max < min) { throw std::invalid_argument("min > max"); }
  return std::min(max, std::max(min, value));
}

// (tagged_object_occupancy_map.hpp:286-288: the ids collected in a set, as a vector)
template <typename Key, typename SetLike>
inline std::vector<Key> GetKeysFromSetLike(const SetLike& set_like)
{
  return std::vector<Key>(set_like.begin(), set_like.end());
}

// (test/voxel_raycasting_test.cpp:17, 88-92: a source of uniform draws from [0, 1))
using UniformUnitRealFunction = std::function<double()>;

// (topology_computation.hpp:335, occupancy_component_map.hpp:267: an optional message sink)
using LoggingFunction = std::function<void(const std::string&)>;

template <typename T, std::memory_order kOrder = std::memory_order_seq_cst>
class CopyableMoveableAtomic
{
public:
  CopyableMoveableAtomic() : value_(T()) {}
  explicit CopyableMoveableAtomic(const T& value) : value_(value) {}
  CopyableMoveableAtomic(const CopyableMoveableAtomic& other) : value_(other.load()) {}
  CopyableMoveableAtomic(CopyableMoveableAtomic&& other) : value_(other.load()) {}
  CopyableMoveableAtomic& operator=(const CopyableMoveableAtomic& other)
  {
    store(other.load());
    return *this;
  }
  CopyableMoveableAtomic& operator=(CopyableMoveableAtomic&& other)
  {
    store(other.load());
    return *this;
  }
  T load() const { return value_.load(kOrder); }
  void store(const T& value) { value_.store(value, kOrder); }
  T fetch_add(const T& operand) { return value_.fetch_add(operand, kOrder); }

private:
  std::atomic<T> value_;
};
}  // namespace utility
}  // namespace common_robotics_utilities
