// TEST INFRASTRUCTURE ONLY -- stand-in that SHADOWS the reference's signed_distance_field.hpp
// (1269 lines of queries/gradients/serialization that need the full common_robotics_utilities)
// with the construction / raw storage / Lock() / min-max surface the SDF generation code uses
// (reference lines 193-211, 724-795, 1234-1264). Put oracle/ref_shim before the reference's
// include directory on the include path.
#pragma once

#include <algorithm>
#include <cstdint>
#include <limits>
#include <string>
#include <vector>

#include <Eigen/Geometry>
#include <common_robotics_utilities/parallelism.hpp>
#include <common_robotics_utilities/voxel_grid.hpp>
#include <voxelized_geometry_tools/vgt_namespace.hpp>

namespace voxelized_geometry_tools
{
VGT_NAMESPACE_BEGIN
template <typename ScalarType>
class SignedDistanceFieldMinimumMaximum
{
public:
  SignedDistanceFieldMinimumMaximum() = default;
  SignedDistanceFieldMinimumMaximum(ScalarType minimum, ScalarType maximum)
      : minimum_(minimum), maximum_(maximum) {}
  ScalarType Minimum() const { return minimum_; }
  ScalarType Maximum() const { return maximum_; }

private:
  ScalarType minimum_ = std::numeric_limits<ScalarType>::infinity();
  ScalarType maximum_ = -std::numeric_limits<ScalarType>::infinity();
};

template <typename ScalarType>
class SignedDistanceField
    : public common_robotics_utilities::voxel_grid::VoxelGridBase<
          ScalarType, std::vector<ScalarType>>
{
public:
  using Base = common_robotics_utilities::voxel_grid::VoxelGridBase<
      ScalarType, std::vector<ScalarType>>;
  SignedDistanceField() = default;
  SignedDistanceField(
      const Eigen::Isometry3d& origin_transform, const std::string& frame,
      const common_robotics_utilities::voxel_grid::VoxelGridSizes& sizes,
      const ScalarType oob_value)
      : Base(origin_transform, sizes, oob_value), frame_(frame) {}

  double Resolution() const { return this->VoxelXSize(); }
  const std::string& Frame() const { return frame_; }
  bool IsLocked() const { return locked_; }

  SignedDistanceFieldMinimumMaximum<ScalarType> GetMinimumMaximum() const
  {
    if (locked_) { return minimum_maximum_; }
    const auto& data = this->GetImmutableRawData();
    if (data.empty()) { return SignedDistanceFieldMinimumMaximum<ScalarType>(); }
    const auto extrema = std::minmax_element(data.begin(), data.end());
    return SignedDistanceFieldMinimumMaximum<ScalarType>(*extrema.first, *extrema.second);
  }
  void Lock()
  {
    minimum_maximum_ = GetMinimumMaximum();
    locked_ = true;
  }
  void Unlock() { locked_ = false; }
  // Adapter hook (not in the reference): install device-computed extrema instead of re-scanning.
  void LockWithKnownExtrema(ScalarType minimum, ScalarType maximum)
  {
    minimum_maximum_ = SignedDistanceFieldMinimumMaximum<ScalarType>(minimum, maximum);
    locked_ = true;
  }

protected:
  bool OnMutableAccess(int64_t, int64_t, int64_t) override { return !locked_; }

private:
  std::string frame_;
  bool locked_ = false;
  SignedDistanceFieldMinimumMaximum<ScalarType> minimum_maximum_;
};

template <typename ScalarType>
class SignedDistanceFieldGenerationParameters
{
public:
  SignedDistanceFieldGenerationParameters()
      : oob_value_(std::numeric_limits<ScalarType>::infinity()) {}
  SignedDistanceFieldGenerationParameters(
      const ScalarType& oob_value,
      const common_robotics_utilities::parallelism::DegreeOfParallelism& parallelism,
      const bool unknown_is_filled, const bool add_virtual_border)
      : oob_value_(oob_value), parallelism_(parallelism), unknown_is_filled_(unknown_is_filled),
        add_virtual_border_(add_virtual_border) {}
  const ScalarType& OOBValue() const { return oob_value_; }
  const common_robotics_utilities::parallelism::DegreeOfParallelism& Parallelism() const
  {
    return parallelism_;
  }
  bool UnknownIsFilled() const { return unknown_is_filled_; }
  bool AddVirtualBorder() const { return add_virtual_border_; }

private:
  ScalarType oob_value_;
  common_robotics_utilities::parallelism::DegreeOfParallelism parallelism_;
  bool unknown_is_filled_ = true;
  bool add_virtual_border_ = false;
};
VGT_NAMESPACE_END
}  // namespace voxelized_geometry_tools
