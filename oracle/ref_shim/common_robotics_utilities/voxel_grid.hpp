// TEST INFRASTRUCTURE ONLY -- stand-in for common_robotics_utilities/voxel_grid.hpp with just the
// surface signed_distance_field_generation.{hpp,cpp} and cpu_pointcloud_voxelization.cpp use
// (index <-> location: floor(p * (1 / voxel_size)) and voxel_size * (index + 0.5), the
// conventions the reference mirrors in-tree at cuda_voxelization_helpers.cu:139-144, 250-255;
// grid extent = voxel count * voxel size). Storage convention as in the real
// library (mirrored in-tree at cuda_voxelization_helpers.cu:281-282): x slowest, z contiguous.
#pragma once

#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <vector>

#include <Eigen/Geometry>
#include <common_robotics_utilities/serialization.hpp>

#ifndef CRU_UNUSED
#define CRU_UNUSED(x) (void)(x)
#endif

namespace common_robotics_utilities
{
namespace voxel_grid
{
struct Vector3i64
{
  int64_t x = 0, y = 0, z = 0;
  Vector3i64() = default;
  Vector3i64(int64_t xi, int64_t yi, int64_t zi) : x(xi), y(yi), z(zi) {}
};

class GridIndex
{
public:
  GridIndex() = default;
  GridIndex(int64_t x, int64_t y, int64_t z) : x_(x), y_(y), z_(z) {}
  const int64_t& X() const { return x_; }
  const int64_t& Y() const { return y_; }
  const int64_t& Z() const { return z_; }
  int64_t& X() { return x_; }
  int64_t& Y() { return y_; }
  int64_t& Z() { return z_; }
  bool operator==(const GridIndex& o) const { return x_ == o.x_ && y_ == o.y_ && z_ == o.z_; }
  bool operator!=(const GridIndex& o) const { return !(*this == o); }

private:
  int64_t x_ = -1, y_ = -1, z_ = -1;
};

class VoxelGridSizes
{
public:
  VoxelGridSizes() = default;
  static VoxelGridSizes FromVoxelCounts(double voxel_size, const Vector3i64& counts)
  {
    if (!(voxel_size > 0.0) || counts.x < 1 || counts.y < 1 || counts.z < 1)
    {
      throw std::invalid_argument("invalid voxel grid sizes");
    }
    VoxelGridSizes sizes;
    sizes.voxel_size_ = voxel_size;
    sizes.counts_ = counts;
    return sizes;
  }
  static VoxelGridSizes FromGridSizes(double voxel_size, const Eigen::Vector3d& grid_sizes)
  {
    return FromVoxelCounts(
        voxel_size,
        Vector3i64(static_cast<int64_t>(std::ceil(grid_sizes(0) / voxel_size)),
                   static_cast<int64_t>(std::ceil(grid_sizes(1) / voxel_size)),
                   static_cast<int64_t>(std::ceil(grid_sizes(2) / voxel_size))));
  }
  int64_t NumXVoxels() const { return counts_.x; }
  int64_t NumYVoxels() const { return counts_.y; }
  int64_t NumZVoxels() const { return counts_.z; }
  int64_t TotalVoxels() const { return counts_.x * counts_.y * counts_.z; }
  double VoxelXSize() const { return voxel_size_; }
  double InvVoxelXSize() const { return 1.0 / voxel_size_; }
  double InverseVoxelXSize() const { return InvVoxelXSize(); }
  Eigen::Vector3d Sizes() const
  {
    return Eigen::Vector3d(static_cast<double>(counts_.x) * voxel_size_,
                           static_cast<double>(counts_.y) * voxel_size_,
                           static_cast<double>(counts_.z) * voxel_size_);
  }
  bool UniformVoxelSize() const { return true; }
  bool operator==(const VoxelGridSizes& o) const
  {
    return voxel_size_ == o.voxel_size_ && counts_.x == o.counts_.x && counts_.y == o.counts_.y
        && counts_.z == o.counts_.z;
  }
  bool operator!=(const VoxelGridSizes& o) const { return !(*this == o); }

private:
  double voxel_size_ = 0.0;
  Vector3i64 counts_;
};

template <typename T>
class GridQuery
{
public:
  GridQuery() = default;
  explicit GridQuery(T* item) : item_(item) {}
  T& Value() const
  {
    if (item_ == nullptr) { throw std::runtime_error("grid query has no value"); }
    return *item_;
  }
  explicit operator bool() const { return item_ != nullptr; }

private:
  T* item_ = nullptr;
};

template <typename T, typename BackingStore = std::vector<T>>
class VoxelGridBase
{
public:
  VoxelGridBase() = default;
  // (the out-of-bounds value of the real class is only ever handed back by accessors the
  // reference's SDF code does not use; kept for the constructors and the serialized form)
  VoxelGridBase(const Eigen::Isometry3d& origin_transform, const VoxelGridSizes& sizes,
                const T& default_value, const T& oob_value)
      : origin_transform_(origin_transform), sizes_(sizes),
        data_(static_cast<size_t>(sizes.TotalVoxels()), default_value), initialized_(true),
        default_value_(default_value), oob_value_(oob_value) {}
  VoxelGridBase(const VoxelGridSizes& sizes, const T& default_value, const T& oob_value)
      : VoxelGridBase(Eigen::Isometry3d::Identity(), sizes, default_value, oob_value) {}
  VoxelGridBase(const Eigen::Isometry3d& origin_transform, const VoxelGridSizes& sizes,
                const T& default_value)
      : VoxelGridBase(origin_transform, sizes, default_value, default_value) {}
  virtual ~VoxelGridBase() {}

  // (signed_distance_field.hpp derives from the grid and implements these)
  using ScalarTypeSerializer = serialization::Serializer<T>;
  using ScalarTypeDeserializer = serialization::Deserializer<T>;
  Eigen::Vector3d VoxelSizes() const
  {
    return Eigen::Vector3d(sizes_.VoxelXSize(), sizes_.VoxelXSize(), sizes_.VoxelXSize());
  }
  Eigen::Vector4d GridIndexToLocationInGridFrame(int64_t x, int64_t y, int64_t z) const
  {
    return GridIndexToLocationInGridFrame(GridIndex(x, y, z));
  }
  GridIndex LocationToGridIndex4d(const Eigen::Vector4d& location) const
  {
    return LocationInGridFrameToGridIndex4d(InverseOriginTransform() * location);
  }
  GridIndex LocationToGridIndex3d(const Eigen::Vector3d& location) const
  {
    return LocationToGridIndex4d(Eigen::Vector4d(location(0), location(1), location(2), 1.0));
  }
  bool CheckLocationInBounds4d(const Eigen::Vector4d& location) const
  {
    return CheckGridIndexInBounds(LocationToGridIndex4d(location));
  }
  bool CheckLocationInBounds(const Eigen::Vector3d& location) const
  {
    return CheckGridIndexInBounds(LocationToGridIndex3d(location));
  }
  bool CheckLocationInBounds(double x, double y, double z) const
  {
    return CheckLocationInBounds4d(Eigen::Vector4d(x, y, z, 1.0));
  }
  GridIndex LocationToGridIndex(double x, double y, double z) const
  {
    return LocationToGridIndex4d(Eigen::Vector4d(x, y, z, 1.0));
  }
  Eigen::Vector4d GridIndexToLocation(const GridIndex& index) const
  {
    return origin_transform_ * GridIndexToLocationInGridFrame(index);
  }
  Eigen::Vector4d GridIndexToLocation(int64_t x, int64_t y, int64_t z) const
  {
    return GridIndexToLocation(GridIndex(x, y, z));
  }
  // The grid's serialized form. PARITY UNPINNED (the real class is not in the reference tree,
  // see serialization.hpp): initialized flag, origin transform, inverse origin transform, the
  // cells as a vector, the sizes (three voxel sizes, three voxel counts), the default and the
  // out-of-bounds value; then whatever the derived class appends.
  uint64_t SerializeSelf(std::vector<uint8_t>& buffer,
                         const ScalarTypeSerializer& value_serializer) const
  {
    const size_t start = buffer.size();
    serialization::SerializeMemcpyable<uint8_t>(static_cast<uint8_t>(initialized_), buffer);
    serialization::SerializeIsometry3d(origin_transform_, buffer);
    serialization::SerializeIsometry3d(origin_transform_.inverse(), buffer);
    serialization::SerializeVectorLike<T, BackingStore>(data_, buffer, value_serializer);
    serialization::SerializeMemcpyable<double>(sizes_.VoxelXSize(), buffer);
    serialization::SerializeMemcpyable<double>(sizes_.VoxelXSize(), buffer);
    serialization::SerializeMemcpyable<double>(sizes_.VoxelXSize(), buffer);
    serialization::SerializeMemcpyable<int64_t>(sizes_.NumXVoxels(), buffer);
    serialization::SerializeMemcpyable<int64_t>(sizes_.NumYVoxels(), buffer);
    serialization::SerializeMemcpyable<int64_t>(sizes_.NumZVoxels(), buffer);
    value_serializer(default_value_, buffer);
    value_serializer(oob_value_, buffer);
    DerivedSerializeSelf(buffer, value_serializer);
    return buffer.size() - start;
  }
  uint64_t DeserializeSelf(const std::vector<uint8_t>& buffer, uint64_t starting_offset,
                           const ScalarTypeDeserializer& value_deserializer)
  {
    uint64_t position = starting_offset;
    const auto initialized = serialization::DeserializeMemcpyable<uint8_t>(buffer, position);
    position += initialized.BytesRead();
    const auto origin = serialization::DeserializeIsometry3d(buffer, position);
    position += origin.BytesRead();
    const auto inverse = serialization::DeserializeIsometry3d(buffer, position);
    position += inverse.BytesRead();
    const auto data =
        serialization::DeserializeVectorLike<T, BackingStore>(buffer, position, value_deserializer);
    position += data.BytesRead();
    double voxel_sizes[3];
    int64_t counts[3];
    for (double& size : voxel_sizes)
    {
      const auto item = serialization::DeserializeMemcpyable<double>(buffer, position);
      size = item.Value();
      position += item.BytesRead();
    }
    for (int64_t& count : counts)
    {
      const auto item = serialization::DeserializeMemcpyable<int64_t>(buffer, position);
      count = item.Value();
      position += item.BytesRead();
    }
    const auto default_value = value_deserializer(buffer, position);
    position += default_value.BytesRead();
    const auto oob_value = value_deserializer(buffer, position);
    position += oob_value.BytesRead();
    if (voxel_sizes[0] != voxel_sizes[1] || voxel_sizes[0] != voxel_sizes[2])
    {
      throw std::invalid_argument("the stand-in grid keeps one voxel size");
    }
    initialized_ = initialized.Value() != 0;
    origin_transform_ = origin.Value();
    sizes_ = VoxelGridSizes::FromVoxelCounts(voxel_sizes[0],
                                             Vector3i64(counts[0], counts[1], counts[2]));
    data_ = data.Value();
    if (static_cast<int64_t>(data_.size()) != sizes_.TotalVoxels())
    {
      throw std::invalid_argument("serialized grid holds the wrong number of cells");
    }
    default_value_ = default_value.Value();
    oob_value_ = oob_value.Value();
    position += DerivedDeserializeSelf(buffer, position, value_deserializer);
    return position - starting_offset;
  }
  const T& DefaultValue() const { return default_value_; }
  const T& OOBValue() const { return oob_value_; }

  bool IsInitialized() const { return initialized_; }
  bool HasUniformVoxelSize() const { return sizes_.UniformVoxelSize(); }
  const Eigen::Isometry3d& OriginTransform() const { return origin_transform_; }
  Eigen::Isometry3d InverseOriginTransform() const { return origin_transform_.inverse(); }
  const VoxelGridSizes& ControlSizes() const { return sizes_; }
  int64_t NumXVoxels() const { return sizes_.NumXVoxels(); }
  int64_t NumYVoxels() const { return sizes_.NumYVoxels(); }
  int64_t NumZVoxels() const { return sizes_.NumZVoxels(); }
  int64_t NumTotalVoxels() const { return sizes_.TotalVoxels(); }
  double VoxelXSize() const { return sizes_.VoxelXSize(); }
  Eigen::Vector3d GridSizes() const { return sizes_.Sizes(); }
  double GridXSize() const { return sizes_.Sizes()(0); }
  double GridYSize() const { return sizes_.Sizes()(1); }
  double GridZSize() const { return sizes_.Sizes()(2); }

  bool CheckGridIndexInBounds(int64_t x, int64_t y, int64_t z) const
  {
    return x >= 0 && x < NumXVoxels() && y >= 0 && y < NumYVoxels() && z >= 0 && z < NumZVoxels();
  }
  bool CheckGridIndexInBounds(const GridIndex& index) const
  {
    return CheckGridIndexInBounds(index.X(), index.Y(), index.Z());
  }
  GridIndex LocationInGridFrameToGridIndex4d(const Eigen::Vector4d& location) const
  {
    const double inverse = sizes_.InvVoxelXSize();
    return GridIndex(static_cast<int64_t>(std::floor(location(0) * inverse)),
                     static_cast<int64_t>(std::floor(location(1) * inverse)),
                     static_cast<int64_t>(std::floor(location(2) * inverse)));
  }
  Eigen::Vector4d GridIndexToLocationInGridFrame(const GridIndex& index) const
  {
    const double voxel = sizes_.VoxelXSize();
    return Eigen::Vector4d(voxel * (static_cast<double>(index.X()) + 0.5),
                           voxel * (static_cast<double>(index.Y()) + 0.5),
                           voxel * (static_cast<double>(index.Z()) + 0.5), 1.0);
  }
  GridQuery<T> GetIndexMutable(const GridIndex& index)
  {
    if (!CheckGridIndexInBounds(index) || !OnMutableAccess(index.X(), index.Y(), index.Z()))
    {
      return GridQuery<T>();
    }
    return GridQuery<T>(&data_[static_cast<size_t>(GetDataIndex(index.X(), index.Y(), index.Z()))]);
  }
  const T& GetDataIndexImmutable(int64_t data_index) const
  {
    return data_.at(static_cast<size_t>(data_index));
  }
  T& GetDataIndexMutable(int64_t data_index) { return data_.at(static_cast<size_t>(data_index)); }
  int64_t GetDataIndex(int64_t x, int64_t y, int64_t z) const
  {
    return (x * NumYVoxels() + y) * NumZVoxels() + z;
  }
  GridQuery<const T> GetIndexImmutable(int64_t x, int64_t y, int64_t z) const
  {
    if (!CheckGridIndexInBounds(x, y, z)) { return GridQuery<const T>(); }
    return GridQuery<const T>(&data_[static_cast<size_t>(GetDataIndex(x, y, z))]);
  }
  GridQuery<const T> GetIndexImmutable(const GridIndex& index) const
  {
    return GetIndexImmutable(index.X(), index.Y(), index.Z());
  }
  bool SetIndex(int64_t x, int64_t y, int64_t z, const T& value)
  {
    if (!CheckGridIndexInBounds(x, y, z) || !OnMutableAccess(x, y, z)) { return false; }
    data_[static_cast<size_t>(GetDataIndex(x, y, z))] = value;
    return true;
  }
  bool SetIndex(const GridIndex& index, const T& value)
  {
    return SetIndex(index.X(), index.Y(), index.Z(), value);
  }
  const BackingStore& GetImmutableRawData() const { return data_; }
  BackingStore& GetMutableRawData() { return data_; }

protected:
  virtual bool OnMutableAccess(int64_t, int64_t, int64_t) { return true; }
  virtual bool OnMutableRawAccess() { return true; }
  virtual std::unique_ptr<VoxelGridBase<T, BackingStore>> DoClone() const { return nullptr; }
  virtual uint64_t DerivedSerializeSelf(std::vector<uint8_t>&, const ScalarTypeSerializer&) const
  {
    return 0;
  }
  virtual uint64_t DerivedDeserializeSelf(const std::vector<uint8_t>&, uint64_t,
                                          const ScalarTypeDeserializer&)
  {
    return 0;
  }

private:
  Eigen::Isometry3d origin_transform_;
  VoxelGridSizes sizes_;
  BackingStore data_;
  bool initialized_ = false;
  T default_value_{};
  T oob_value_{};
};

template <typename T>
using VoxelGrid = VoxelGridBase<T, std::vector<T>>;
}  // namespace voxel_grid
}  // namespace common_robotics_utilities

namespace std
{
template <>
struct hash<common_robotics_utilities::voxel_grid::GridIndex>
{
  size_t operator()(const common_robotics_utilities::voxel_grid::GridIndex& index) const
  {
    // (any hash will do: the reference only uses the map as a set of visited cells)
    size_t h = std::hash<int64_t>()(index.X());
    h = h * 1000003u + std::hash<int64_t>()(index.Y());
    return h * 1000003u + std::hash<int64_t>()(index.Z());
  }
};
}  // namespace std
