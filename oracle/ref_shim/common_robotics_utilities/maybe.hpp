// TEST INFRASTRUCTURE ONLY -- stand-in for common_robotics_utilities/maybe.hpp: an optional that
// throws when read empty (the surface signed_distance_field.hpp uses for its query results).
#pragma once

#include <stdexcept>
#include <utility>

namespace common_robotics_utilities
{
template <typename T>
class OwningMaybe
{
public:
  OwningMaybe() = default;
  explicit OwningMaybe(const T& value) : value_(value), has_value_(true) {}
  void Reset() { has_value_ = false; }
  const T& Value() const
  {
    if (!has_value_) { throw std::runtime_error("OwningMaybe does not have value"); }
    return value_;
  }
  bool HasValue() const { return has_value_; }
  explicit operator bool() const { return has_value_; }

private:
  T value_{};
  bool has_value_ = false;
};
}  // namespace common_robotics_utilities
