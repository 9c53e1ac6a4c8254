// TEST INFRASTRUCTURE ONLY -- stand-in for common_robotics_utilities/zlib_helpers.hpp: named by
// signed_distance_field.hpp's file members, never called by the oracle; every function throws.
#pragma once

#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace common_robotics_utilities
{
namespace zlib_helpers
{
inline std::vector<uint8_t> CompressBytes(const std::vector<uint8_t>&)
{
  throw std::runtime_error("file formats are not part of the oracle");
}
inline std::vector<uint8_t> DecompressBytes(const std::vector<uint8_t>&)
{
  throw std::runtime_error("file formats are not part of the oracle");
}
inline std::vector<uint8_t> LoadFromFileAndDecompress(const std::string&)
{
  throw std::runtime_error("file formats are not part of the oracle");
}
inline void CompressAndWriteToFile(const std::vector<uint8_t>&, const std::string&)
{
  throw std::runtime_error("file formats are not part of the oracle");
}
}  // namespace zlib_helpers
}  // namespace common_robotics_utilities
