// TEST INFRASTRUCTURE ONLY -- stand-in for common_robotics_utilities/zlib_helpers.hpp (not in the
// reference tree, unpinned): CompressBytes is one zlib stream of the whole buffer at the
// library's default level, DecompressBytes inflates one zlib stream of unknown size. Used by the
// reference's own SaveToFile / LoadFromFile (signed_distance_field.hpp:643-722) in the oracle
// build. PARITY UNPINNED at this layer (see serialization.hpp).
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include <zlib.h>

namespace common_robotics_utilities
{
namespace zlib_helpers
{
inline std::vector<uint8_t> CompressBytes(const std::vector<uint8_t>& uncompressed)
{
  uLongf capacity = compressBound(static_cast<uLong>(uncompressed.size()));
  std::vector<uint8_t> compressed(capacity);
  if (compress2(compressed.data(), &capacity, uncompressed.data(),
                static_cast<uLong>(uncompressed.size()), Z_DEFAULT_COMPRESSION) != Z_OK)
  {
    throw std::runtime_error("ZLIB compression failed");
  }
  compressed.resize(capacity);
  return compressed;
}

inline std::vector<uint8_t> DecompressBytes(const std::vector<uint8_t>& compressed)
{
  z_stream stream{};
  if (inflateInit(&stream) != Z_OK)
  {
    throw std::runtime_error("ZLIB unable to init inflate stream");
  }
  stream.next_in = const_cast<Bytef*>(compressed.data());
  stream.avail_in = static_cast<uInt>(compressed.size());
  std::vector<uint8_t> decompressed;
  std::vector<uint8_t> chunk(1 << 20);
  int status = Z_OK;
  while (status != Z_STREAM_END)
  {
    stream.next_out = chunk.data();
    stream.avail_out = static_cast<uInt>(chunk.size());
    status = inflate(&stream, Z_NO_FLUSH);
    if (status != Z_OK && status != Z_STREAM_END)
    {
      inflateEnd(&stream);
      throw std::runtime_error("ZLIB decompression failed");
    }
    decompressed.insert(decompressed.end(), chunk.data(),
                        chunk.data() + (chunk.size() - stream.avail_out));
  }
  inflateEnd(&stream);
  return decompressed;
}
}  // namespace zlib_helpers
}  // namespace common_robotics_utilities
