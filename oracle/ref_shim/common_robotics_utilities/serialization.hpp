// TEST INFRASTRUCTURE ONLY -- stand-in for common_robotics_utilities/serialization.hpp, so that
// the reference's own SaveToFile / LoadFromFile / Serialize / Deserialize members
// (signed_distance_field.hpp:551-596, 622-722) run here as written.
//
// PARITY UNPINNED at this layer: common_robotics_utilities is not in the reference tree and is
// not pinned by it (package.xml.ros2:12). The byte layout below restates that library's
// published serializer as this repo's author knows it: every item is appended as its raw
// little-endian bytes (SerializeMemcpyable), a string or vector is a uint64 element count
// followed by its elements (SerializeString / SerializeVectorLike), an Isometry3d is the sixteen
// doubles of its 4x4 matrix in column-major order (SerializeIsometry3d). What the reference build
// pins is everything the reference tree itself owns: the file magics, the compress switch, the
// order of the derived members (frame, then the locked byte) after the grid's own bytes.
#pragma once

#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <Eigen/Geometry>

namespace common_robotics_utilities
{
namespace serialization
{
template <typename T>
using Serializer = std::function<uint64_t(const T&, std::vector<uint8_t>&)>;

template <typename T>
class Deserialized
{
public:
  Deserialized() = default;
  Deserialized(const T& value, uint64_t bytes_read) : value_(value), bytes_read_(bytes_read) {}
  const T& Value() const { return value_; }
  uint64_t BytesRead() const { return bytes_read_; }

private:
  T value_{};
  uint64_t bytes_read_ = 0;
};

template <typename T>
using Deserializer = std::function<Deserialized<T>(const std::vector<uint8_t>&, uint64_t)>;

template <typename T>
inline Deserialized<T> MakeDeserialized(const T& value, uint64_t bytes_read)
{
  return Deserialized<T>(value, bytes_read);
}

template <typename T>
inline uint64_t SerializeMemcpyable(const T& item, std::vector<uint8_t>& buffer)
{
  const size_t start = buffer.size();
  buffer.resize(start + sizeof(T));
  std::memcpy(buffer.data() + start, &item, sizeof(T));
  return sizeof(T);
}

template <typename T>
inline Deserialized<T> DeserializeMemcpyable(const std::vector<uint8_t>& buffer, uint64_t offset)
{
  if (offset + sizeof(T) > buffer.size())
  {
    throw std::invalid_argument("Not enough room in the provided buffer");
  }
  T item;
  std::memcpy(&item, buffer.data() + offset, sizeof(T));
  return Deserialized<T>(item, sizeof(T));
}

template <typename T, typename Container = std::vector<T>>
inline uint64_t SerializeVectorLike(const Container& items, std::vector<uint8_t>& buffer,
                                    const Serializer<T>& item_serializer)
{
  const size_t start = buffer.size();
  SerializeMemcpyable<uint64_t>(static_cast<uint64_t>(items.size()), buffer);
  for (const T& item : items)
  {
    item_serializer(item, buffer);
  }
  return buffer.size() - start;
}

template <typename T, typename Container = std::vector<T>>
inline Deserialized<Container> DeserializeVectorLike(const std::vector<uint8_t>& buffer,
                                                     uint64_t offset,
                                                     const Deserializer<T>& item_deserializer)
{
  uint64_t position = offset;
  const auto count = DeserializeMemcpyable<uint64_t>(buffer, position);
  position += count.BytesRead();
  Container items;
  items.reserve(static_cast<size_t>(count.Value()));
  for (uint64_t i = 0; i < count.Value(); i++)
  {
    const auto item = item_deserializer(buffer, position);
    items.push_back(item.Value());
    position += item.BytesRead();
  }
  return Deserialized<Container>(items, position - offset);
}

template <typename CharT = char>
inline uint64_t SerializeString(const std::basic_string<CharT>& text, std::vector<uint8_t>& buffer)
{
  return SerializeVectorLike<CharT, std::basic_string<CharT>>(text, buffer,
                                                              SerializeMemcpyable<CharT>);
}

template <typename CharT = char>
inline Deserialized<std::basic_string<CharT>> DeserializeString(
    const std::vector<uint8_t>& buffer, uint64_t offset)
{
  return DeserializeVectorLike<CharT, std::basic_string<CharT>>(buffer, offset,
                                                                DeserializeMemcpyable<CharT>);
}

inline uint64_t SerializeIsometry3d(const Eigen::Isometry3d& transform, std::vector<uint8_t>& buffer)
{
  const size_t start = buffer.size();
  buffer.resize(start + sizeof(double) * 16);
  std::memcpy(buffer.data() + start, transform.data(), sizeof(double) * 16);
  return sizeof(double) * 16;
}

inline Deserialized<Eigen::Isometry3d> DeserializeIsometry3d(const std::vector<uint8_t>& buffer,
                                                             uint64_t offset)
{
  if (offset + sizeof(double) * 16 > buffer.size())
  {
    throw std::invalid_argument("Not enough room in the provided buffer");
  }
  Eigen::Isometry3d transform;
  std::memcpy(transform.data(), buffer.data() + offset, sizeof(double) * 16);
  return Deserialized<Eigen::Isometry3d>(transform, sizeof(double) * 16);
}
}  // namespace serialization
}  // namespace common_robotics_utilities
