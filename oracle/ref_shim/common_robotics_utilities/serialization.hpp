// TEST INFRASTRUCTURE ONLY -- stand-in for common_robotics_utilities/serialization.hpp: the
// names signed_distance_field.hpp mentions in its (de)serialization members, which the oracle
// never calls (file formats are out of scope); every function throws.
#pragma once

#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

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
inline uint64_t SerializeMemcpyable(const T&, std::vector<uint8_t>&)
{
  throw std::runtime_error("serialization is not part of the oracle");
}

template <typename T>
inline Deserialized<T> DeserializeMemcpyable(const std::vector<uint8_t>&, uint64_t)
{
  throw std::runtime_error("serialization is not part of the oracle");
}

template <typename CharT = char>
inline uint64_t SerializeString(const std::basic_string<CharT>&, std::vector<uint8_t>&)
{
  throw std::runtime_error("serialization is not part of the oracle");
}

template <typename CharT = char>
inline Deserialized<std::basic_string<CharT>> DeserializeString(
    const std::vector<uint8_t>&, uint64_t)
{
  throw std::runtime_error("serialization is not part of the oracle");
}
}  // namespace serialization
}  // namespace common_robotics_utilities
