// On-disk formats of the grids on this path, so that results computed
// on the device round-trip through the reference's own files:
//   SignedDistanceField<T>::SaveToFile / LoadFromFile   "SDFZ" / "SDFR"
//       include/voxelized_geometry_tools/signed_distance_field.hpp:622-722, derived members :551-596
//   OccupancyMap::SaveToFile / LoadFromFile             "CMGZ" / "CMGR"
//       src/voxelized_geometry_tools/occupancy_map.cpp:100-193, derived members :56-84,
//       cells :23-46 (one float)
// A file is a four-byte magic followed by the serialized grid, raw ("..R") or as one zlib stream
// ("..Z"). The serialized grid is common_robotics_utilities' VoxelGridBase form followed by the
// derived class's members. That library is not part of the reference tree and is not pinned by
// it, so its layout is restated here (and stated as "parity unpinned" in DESIGN.md):
//   u8   initialized
//   f64  origin transform, 4x4 column-major          (SerializeIsometry3d)
//   f64  inverse origin transform, 4x4 column-major
//   u64  cell count, then the cells                   (SerializeVectorLike; x slowest, z fastest)
//   f64  voxel size x, y, z;  i64 voxel count x, y, z (the grid sizes)
//   cell default value;  cell out-of-bounds value
//   derived: u64 length + bytes of the frame name; SDF only: u8 locked
// Host code only (zlib for the compressed forms); the _dev entry copies a device-resident grid
// out first.
#include "common.cuh"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <zlib.h>

namespace vgt_b200
{
namespace
{
constexpr size_t kHeaderBytes = 4;

size_t CellBytes(int kind)
{
  return kind == VGT_B200_GRID_FILE_SDF_F64 ? 8 : 4;
}

const char* Magic(int kind, bool compressed)
{
  if (kind == VGT_B200_GRID_FILE_OCCUPANCY)
  {
    return compressed ? "CMGZ" : "CMGR";
  }
  return compressed ? "SDFZ" : "SDFR";
}

template <typename T>
void Append(std::vector<uint8_t>& buffer, const T& item)
{
  const size_t start = buffer.size();
  buffer.resize(start + sizeof(T));
  std::memcpy(buffer.data() + start, &item, sizeof(T));
}

void AppendCellValue(std::vector<uint8_t>& buffer, int kind, double value)
{
  if (CellBytes(kind) == 8)
  {
    Append<double>(buffer, value);
  }
  else
  {
    Append<float>(buffer, static_cast<float>(value));
  }
}

// Rigid inverse: R^T and -(R^T t), each dot product summed left to right.
void InvertRigid(const double* m, double* out)
{
  std::memset(out, 0, sizeof(double) * 16);
  for (int r = 0; r < 3; r++)
  {
    for (int c = 0; c < 3; c++)
    {
      out[c * 4 + r] = m[r * 4 + c];
    }
  }
  for (int r = 0; r < 3; r++)
  {
    out[12 + r] = -(out[0 * 4 + r] * m[12] + out[1 * 4 + r] * m[13] + out[2 * 4 + r] * m[14]);
  }
  out[15] = 1.0;
}

class Reader
{
public:
  Reader(const uint8_t* data, size_t size) : data_(data), size_(size) {}
  template <typename T>
  bool Take(T* item)
  {
    if (position_ + sizeof(T) > size_)
    {
      return false;
    }
    std::memcpy(item, data_ + position_, sizeof(T));
    position_ += sizeof(T);
    return true;
  }
  bool Skip(size_t bytes)
  {
    if (position_ + bytes > size_ || position_ + bytes < position_)
    {
      return false;
    }
    position_ += bytes;
    return true;
  }
  const uint8_t* Here() const { return data_ + position_; }
  size_t Position() const { return position_; }

private:
  const uint8_t* data_;
  size_t size_;
  size_t position_ = 0;
};

bool TakeCellValue(Reader& reader, int kind, double* value)
{
  if (CellBytes(kind) == 8)
  {
    return reader.Take<double>(value);
  }
  float narrow = 0.0f;
  const bool ok = reader.Take<float>(&narrow);
  *value = static_cast<double>(narrow);
  return ok;
}

int ReadWholeFile(const char* path, int kind, std::vector<uint8_t>* payload)
{
  // (the messages are the reference's: signed_distance_field.hpp:676-721, occupancy_map.cpp:147-192)
  FILE* file = std::fopen(path, "rb");
  if (file == nullptr)
  {
    return FailInvalid("File does not exist");
  }
  std::fseek(file, 0, SEEK_END);
  const long size = std::ftell(file);
  std::fseek(file, 0, SEEK_SET);
  if (size < static_cast<long>(kHeaderBytes))
  {
    std::fclose(file);
    return FailInvalid("File is too small");
  }
  char magic[kHeaderBytes + 1] = {0, 0, 0, 0, 0};
  std::vector<uint8_t> body(static_cast<size_t>(size) - kHeaderBytes);
  const bool read_ok = std::fread(magic, 1, kHeaderBytes, file) == kHeaderBytes
      && (body.empty() || std::fread(body.data(), 1, body.size(), file) == body.size());
  std::fclose(file);
  if (!read_ok)
  {
    return FailInvalid("File could not be read");
  }
  if (std::strcmp(magic, Magic(kind, false)) == 0)
  {
    payload->swap(body);
    return VGT_B200_OK;
  }
  if (std::strcmp(magic, Magic(kind, true)) != 0)
  {
    return FailInvalid("File has invalid header [%s]", magic);
  }
  z_stream stream{};
  if (inflateInit(&stream) != Z_OK)
  {
    return FailInvalid("zlib: inflateInit failed");
  }
  payload->clear();
  std::vector<uint8_t> chunk(size_t{1} << 22);
  size_t consumed = 0;
  int status = Z_OK;
  while (status != Z_STREAM_END)
  {
    if (stream.avail_in == 0 && consumed < body.size())
    {
      // (avail_in is 32 bits: feed the body in pieces)
      const size_t piece = std::min<size_t>(body.size() - consumed, size_t{1} << 30);
      stream.next_in = body.data() + consumed;
      stream.avail_in = static_cast<uInt>(piece);
      consumed += piece;
    }
    stream.next_out = chunk.data();
    stream.avail_out = static_cast<uInt>(chunk.size());
    status = inflate(&stream, Z_NO_FLUSH);
    if (status != Z_OK && status != Z_STREAM_END)
    {
      inflateEnd(&stream);
      return FailInvalid("zlib: the compressed grid is damaged (inflate returned %d)", status);
    }
    payload->insert(payload->end(), chunk.data(), chunk.data() + (chunk.size() - stream.avail_out));
  }
  inflateEnd(&stream);
  return VGT_B200_OK;
}

// Parses everything but the cells; *cells points at them inside the payload.
int ParsePayload(const std::vector<uint8_t>& payload, int kind, vgt_b200_grid_file_info* info,
                 char* frame, int64_t frame_capacity, const uint8_t** cells)
{
  Reader reader(payload.data(), payload.size());
  uint8_t initialized = 0;
  uint64_t cell_count = 0;
  bool ok = reader.Take(&initialized);
  for (int i = 0; ok && i < 16; i++) { ok = reader.Take(&info->origin_transform[i]); }
  for (int i = 0; ok && i < 16; i++) { ok = reader.Take(&info->inverse_origin_transform[i]); }
  ok = ok && reader.Take(&cell_count);
  *cells = reader.Here();
  ok = ok && cell_count <= payload.size() && reader.Skip(static_cast<size_t>(cell_count) * CellBytes(kind));
  for (int i = 0; ok && i < 3; i++) { ok = reader.Take(&info->voxel_size[i]); }
  int64_t counts[3] = {0, 0, 0};
  for (int i = 0; ok && i < 3; i++) { ok = reader.Take(&counts[i]); }
  ok = ok && TakeCellValue(reader, kind, &info->default_value)
      && TakeCellValue(reader, kind, &info->oob_value);
  uint64_t frame_length = 0;
  ok = ok && reader.Take(&frame_length);
  const uint8_t* frame_bytes = reader.Here();
  ok = ok && frame_length <= payload.size() && reader.Skip(static_cast<size_t>(frame_length));
  uint8_t locked = 0;
  if (kind != VGT_B200_GRID_FILE_OCCUPANCY)
  {
    ok = ok && reader.Take(&locked);
  }
  if (!ok)
  {
    return FailInvalid("Not enough room in the provided buffer");
  }
  if (counts[0] < 0 || counts[1] < 0 || counts[2] < 0
      || static_cast<uint64_t>(counts[0]) * static_cast<uint64_t>(counts[1])
              * static_cast<uint64_t>(counts[2])
          != cell_count)
  {
    return FailInvalid("serialized grid holds %llu cells for %lld x %lld x %lld voxels",
                       static_cast<unsigned long long>(cell_count), static_cast<long long>(counts[0]),
                       static_cast<long long>(counts[1]), static_cast<long long>(counts[2]));
  }
  info->nx = counts[0];
  info->ny = counts[1];
  info->nz = counts[2];
  info->initialized = initialized;
  info->locked = locked;
  info->frame_length = static_cast<int64_t>(frame_length);
  info->payload_bytes = static_cast<int64_t>(payload.size());
  if (frame != nullptr && frame_capacity > 0)
  {
    const size_t copied = std::min<size_t>(static_cast<size_t>(frame_length),
                                           static_cast<size_t>(frame_capacity) - 1);
    std::memcpy(frame, frame_bytes, copied);
    frame[copied] = 0;
  }
  return VGT_B200_OK;
}

int SaveGrid(const char* path, int kind, int compress, const void* cells,
             const vgt_b200_grid_file_info* info, const char* frame)
{
  if (path == nullptr || info == nullptr || (cells == nullptr && info->nx * info->ny * info->nz > 0))
  {
    return FailInvalid("grid file: null argument");
  }
  if (kind < VGT_B200_GRID_FILE_SDF_F32 || kind > VGT_B200_GRID_FILE_OCCUPANCY)
  {
    return FailInvalid("grid file: unknown kind %d", kind);
  }
  if (info->nx < 0 || info->ny < 0 || info->nz < 0)
  {
    return FailInvalid("grid file: negative voxel count");
  }
  // (both map types refuse non-uniform voxels: sdf.hpp:612-620, occupancy_map.cpp:72-73)
  if (info->voxel_size[0] != info->voxel_size[1] || info->voxel_size[0] != info->voxel_size[2])
  {
    return FailInvalid("Signed distance field cannot have non-uniform voxel sizes");
  }
  const uint64_t cell_count = static_cast<uint64_t>(info->nx) * static_cast<uint64_t>(info->ny)
      * static_cast<uint64_t>(info->nz);
  const size_t cell_bytes = static_cast<size_t>(cell_count) * CellBytes(kind);
  const std::string frame_name(frame != nullptr ? frame : "");
  std::vector<uint8_t> buffer;
  buffer.reserve(cell_bytes + 512 + frame_name.size());
  Append<uint8_t>(buffer, info->initialized != 0 ? 1 : 0);
  double inverse[16];
  InvertRigid(info->origin_transform, inverse);
  for (int i = 0; i < 16; i++) { Append<double>(buffer, info->origin_transform[i]); }
  for (int i = 0; i < 16; i++) { Append<double>(buffer, inverse[i]); }
  Append<uint64_t>(buffer, cell_count);
  const size_t cells_at = buffer.size();
  buffer.resize(cells_at + cell_bytes);
  if (cell_bytes > 0)
  {
    std::memcpy(buffer.data() + cells_at, cells, cell_bytes);
  }
  for (int i = 0; i < 3; i++) { Append<double>(buffer, info->voxel_size[i]); }
  Append<int64_t>(buffer, info->nx);
  Append<int64_t>(buffer, info->ny);
  Append<int64_t>(buffer, info->nz);
  AppendCellValue(buffer, kind, info->default_value);
  AppendCellValue(buffer, kind, info->oob_value);
  Append<uint64_t>(buffer, static_cast<uint64_t>(frame_name.size()));
  buffer.insert(buffer.end(), frame_name.begin(), frame_name.end());
  if (kind != VGT_B200_GRID_FILE_OCCUPANCY)
  {
    Append<uint8_t>(buffer, info->locked != 0 ? 1 : 0);
  }

  FILE* file = std::fopen(path, "wb");
  if (file == nullptr)
  {
    return FailInvalid("grid file: cannot open [%s] for writing", path);
  }
  bool ok = std::fwrite(Magic(kind, compress != 0), 1, kHeaderBytes, file) == kHeaderBytes;
  if (compress == 0)
  {
    ok = ok && std::fwrite(buffer.data(), 1, buffer.size(), file) == buffer.size();
  }
  else
  {
    // one zlib stream at the library's default level (what compress2 / CompressBytes produce),
    // fed in pieces because zlib counts input in 32 bits
    z_stream stream{};
    if (deflateInit(&stream, Z_DEFAULT_COMPRESSION) != Z_OK)
    {
      std::fclose(file);
      return FailInvalid("zlib: deflateInit failed");
    }
    std::vector<uint8_t> chunk(size_t{1} << 22);
    size_t consumed = 0;
    int status = Z_OK;
    while (ok && status != Z_STREAM_END)
    {
      if (stream.avail_in == 0 && consumed < buffer.size())
      {
        const size_t piece = std::min<size_t>(buffer.size() - consumed, size_t{1} << 30);
        stream.next_in = buffer.data() + consumed;
        stream.avail_in = static_cast<uInt>(piece);
        consumed += piece;
      }
      stream.next_out = chunk.data();
      stream.avail_out = static_cast<uInt>(chunk.size());
      status = deflate(&stream, consumed == buffer.size() ? Z_FINISH : Z_NO_FLUSH);
      if (status != Z_OK && status != Z_STREAM_END && status != Z_BUF_ERROR)
      {
        ok = false;
        break;
      }
      const size_t produced = chunk.size() - stream.avail_out;
      ok = produced == 0 || std::fwrite(chunk.data(), 1, produced, file) == produced;
    }
    deflateEnd(&stream);
  }
  ok = (std::fclose(file) == 0) && ok;
  return ok ? VGT_B200_OK : FailInvalid("grid file: writing [%s] failed", path);
}
}  // namespace
}  // namespace vgt_b200

extern "C"
{
int vgt_b200_grid_file_save(const char* path, int kind, int compress, const void* cells,
                            const vgt_b200_grid_file_info* info, const char* frame)
{
  return vgt_b200::SaveGrid(path, kind, compress, cells, info, frame);
}

int vgt_b200_grid_file_save_dev(const char* path, int kind, int compress, const void* d_cells,
                                const vgt_b200_grid_file_info* info, const char* frame, int device,
                                void* stream)
{
  using namespace vgt_b200;
  if (info == nullptr || d_cells == nullptr)
  {
    return FailInvalid("grid file: null argument");
  }
  if (kind < VGT_B200_GRID_FILE_SDF_F32 || kind > VGT_B200_GRID_FILE_OCCUPANCY || info->nx < 0
      || info->ny < 0 || info->nz < 0)
  {
    return FailInvalid("grid file: unknown kind or negative voxel count");
  }
  VGT_CUDA_TRY(cudaSetDevice(device), "cudaSetDevice");
  const size_t bytes = static_cast<size_t>(info->nx) * static_cast<size_t>(info->ny)
      * static_cast<size_t>(info->nz) * CellBytes(kind);
  std::vector<uint8_t> host(bytes);
  cudaStream_t cuda_stream = static_cast<cudaStream_t>(stream);
  VGT_CUDA_TRY(cudaMemcpyAsync(host.data(), d_cells, bytes, cudaMemcpyDeviceToHost, cuda_stream),
               "grid file copy-out");
  VGT_CUDA_TRY(cudaStreamSynchronize(cuda_stream), "grid file copy-out");
  return SaveGrid(path, kind, compress, host.data(), info, frame);
}

int vgt_b200_grid_file_probe(const char* path, int kind, vgt_b200_grid_file_info* info,
                             char* frame, int64_t frame_capacity)
{
  using namespace vgt_b200;
  if (path == nullptr || info == nullptr
      || kind < VGT_B200_GRID_FILE_SDF_F32 || kind > VGT_B200_GRID_FILE_OCCUPANCY)
  {
    return FailInvalid("grid file: null argument or unknown kind");
  }
  std::vector<uint8_t> payload;
  const int status = ReadWholeFile(path, kind, &payload);
  if (status != VGT_B200_OK)
  {
    return status;
  }
  const uint8_t* cells = nullptr;
  return ParsePayload(payload, kind, info, frame, frame_capacity, &cells);
}

int vgt_b200_grid_file_load(const char* path, int kind, void* cells, int64_t cell_capacity,
                            vgt_b200_grid_file_info* info, char* frame, int64_t frame_capacity)
{
  using namespace vgt_b200;
  if (path == nullptr || info == nullptr
      || kind < VGT_B200_GRID_FILE_SDF_F32 || kind > VGT_B200_GRID_FILE_OCCUPANCY)
  {
    return FailInvalid("grid file: null argument or unknown kind");
  }
  std::vector<uint8_t> payload;
  int status = ReadWholeFile(path, kind, &payload);
  if (status != VGT_B200_OK)
  {
    return status;
  }
  const uint8_t* stored = nullptr;
  status = ParsePayload(payload, kind, info, frame, frame_capacity, &stored);
  if (status != VGT_B200_OK)
  {
    return status;
  }
  const int64_t count = info->nx * info->ny * info->nz;
  if (count > cell_capacity || (cells == nullptr && count > 0))
  {
    return FailInvalid("grid file: %lld cells do not fit the caller's buffer of %lld",
                       static_cast<long long>(count), static_cast<long long>(cell_capacity));
  }
  if (count > 0)
  {
    std::memcpy(cells, stored, static_cast<size_t>(count) * CellBytes(kind));
  }
  return VGT_B200_OK;
}
}  // extern "C"
