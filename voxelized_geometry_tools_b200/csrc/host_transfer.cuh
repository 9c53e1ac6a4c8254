// Host <-> device transfers for the host-pointer entry points when the caller's buffers are
// PAGEABLE (std::vector storage of the reference's grids, numpy arrays): the driver's own staging
// of pageable memory runs at ~11 GB/s host-to-device and ~21 GB/s device-to-host on the B200
// hosts (profiles/host_copy_bandwidth.py), a quarter of what the PCIe link gives pinned memory.
// Here the copy goes through a small ring of pinned slots, filled / drained by a few host threads
// (a multi-threaded memcpy reaches 50-75 GB/s on the same host), so the DMA of one slot overlaps
// the memcpy of the next. Pinned or registered buffers bypass all of this (plain async copies).
//
// Host-only code; no reference counterpart (the reference's device helper uses blocking
// cudaMemcpy on pageable vectors, cuda_voxelization_helpers.cu:676-680, 754-767).
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace vgt_b200
{
// ------------------------------------------------------------------------------------------------
// A few persistent worker threads that copy ranges of host memory.
// ------------------------------------------------------------------------------------------------
class HostCopyPool
{
public:
  static HostCopyPool& Instance()
  {
    static HostCopyPool pool;
    return pool;
  }

  // The same copy on the calling thread alone (callers that are themselves one of several
  // parallel threads, e.g. the per-device threads of the multi-device entry: the shared workers
  // serve one caller at a time and would serialise them).
  static void Copy2DInline(char* dst, size_t dst_pitch, const char* src, size_t src_pitch,
                           size_t row_bytes, size_t rows)
  {
    CopyRange(dst, dst_pitch, src, src_pitch, row_bytes, rows, 0, row_bytes * rows);
  }

  // rows x row_bytes from src (pitch src_pitch) to dst (pitch dst_pitch), split over the workers.
  // Blocks until done.
  void Copy2D(char* dst, size_t dst_pitch, const char* src, size_t src_pitch, size_t row_bytes,
              size_t rows)
  {
    const size_t total = row_bytes * rows;
    if (total == 0)
    {
      return;
    }
    const size_t parts = std::min<size_t>(workers_.size() + 1,
                                          std::max<size_t>(1, total / (size_t{1} << 20)));
    if (parts <= 1)
    {
      CopyRange(dst, dst_pitch, src, src_pitch, row_bytes, rows, 0, total);
      return;
    }
    std::unique_lock<std::mutex> call_lock(call_mutex_);  // one caller at a time
    {
      std::lock_guard<std::mutex> lock(mutex_);
      job_ = Job{dst, dst_pitch, src, src_pitch, row_bytes, rows, total, parts};
      next_part_ = 0;
      remaining_ = parts;
      generation_++;
    }
    wake_.notify_all();
    RunParts();  // the calling thread works too
    std::unique_lock<std::mutex> lock(mutex_);
    done_.wait(lock, [&] { return remaining_ == 0; });
  }

private:
  struct Job
  {
    char* dst;
    size_t dst_pitch;
    const char* src;
    size_t src_pitch;
    size_t row_bytes;
    size_t rows;
    size_t total;
    size_t parts;
  };

  HostCopyPool()
  {
    const unsigned hardware = std::max(2u, std::thread::hardware_concurrency());
    const unsigned count = std::min(7u, hardware / 2);
    for (unsigned i = 0; i < count; i++)
    {
      workers_.emplace_back([this] { WorkerLoop(); });
    }
  }

  ~HostCopyPool()
  {
    {
      std::lock_guard<std::mutex> lock(mutex_);
      stop_ = true;
    }
    wake_.notify_all();
    for (auto& worker : workers_)
    {
      worker.join();
    }
  }

  // bytes [begin, end) of the logical rows x row_bytes array
  static void CopyRange(char* dst, size_t dst_pitch, const char* src, size_t src_pitch,
                        size_t row_bytes, size_t rows, size_t begin, size_t end)
  {
    if (dst_pitch == row_bytes && src_pitch == row_bytes)
    {
      std::memcpy(dst + begin, src + begin, end - begin);
      return;
    }
    (void)rows;
    while (begin < end)
    {
      const size_t row = begin / row_bytes;
      const size_t offset = begin - row * row_bytes;
      const size_t chunk = std::min(end - begin, row_bytes - offset);
      std::memcpy(dst + row * dst_pitch + offset, src + row * src_pitch + offset, chunk);
      begin += chunk;
    }
  }

  void RunParts()
  {
    while (true)
    {
      Job job;
      size_t part;
      {
        std::lock_guard<std::mutex> lock(mutex_);
        if (next_part_ >= job_.parts)
        {
          return;
        }
        job = job_;
        part = next_part_++;
      }
      // 64-byte aligned part boundaries
      const size_t per_part = ((job.total + job.parts - 1) / job.parts + 63) & ~size_t{63};
      const size_t begin = std::min(job.total, part * per_part);
      const size_t end = std::min(job.total, begin + per_part);
      CopyRange(job.dst, job.dst_pitch, job.src, job.src_pitch, job.row_bytes, job.rows, begin,
                end);
      {
        std::lock_guard<std::mutex> lock(mutex_);
        remaining_--;
        if (remaining_ == 0)
        {
          done_.notify_all();
        }
      }
    }
  }

  void WorkerLoop()
  {
    uint64_t seen = 0;
    while (true)
    {
      {
        std::unique_lock<std::mutex> lock(mutex_);
        wake_.wait(lock, [&] { return stop_ || generation_ != seen; });
        if (stop_)
        {
          return;
        }
        seen = generation_;
      }
      RunParts();
    }
  }

  std::vector<std::thread> workers_;
  std::mutex call_mutex_;
  std::mutex mutex_;
  std::condition_variable wake_;
  std::condition_variable done_;
  Job job_{};
  size_t next_part_ = 0;
  size_t remaining_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

// ------------------------------------------------------------------------------------------------
// Pinned staging slots, kept for the life of the process (allocating pinned memory is slow).
// ------------------------------------------------------------------------------------------------
constexpr size_t kStagingSlotBytes = size_t{32} << 20;
constexpr int kStagingSlotsPerCall = 3;

class StagingSlots
{
public:
  // Takes kStagingSlotsPerCall pinned slots out of the process-wide cache (allocating the first
  // time); returns them on destruction. ok() is false when pinned memory cannot be allocated,
  // in which case the caller falls back to plain copies.
  StagingSlots()
  {
    std::lock_guard<std::mutex> lock(CacheMutex());
    auto& cache = Cache();
    for (int i = 0; i < kStagingSlotsPerCall; i++)
    {
      void* slot = nullptr;
      if (!cache.empty())
      {
        slot = cache.back();
        cache.pop_back();
      }
      else if (cudaHostAlloc(&slot, kStagingSlotBytes, cudaHostAllocPortable) != cudaSuccess)
      {
        cudaGetLastError();
        slot = nullptr;
      }
      if (slot == nullptr)
      {
        ok_ = false;
        break;
      }
      slots_[count_] = static_cast<char*>(slot);
      cudaEventCreateWithFlags(&events_[count_], cudaEventDisableTiming);
      count_++;
    }
  }
  ~StagingSlots()
  {
    std::lock_guard<std::mutex> lock(CacheMutex());
    for (int i = 0; i < count_; i++)
    {
      cudaEventDestroy(events_[i]);
      Cache().push_back(slots_[i]);
    }
  }
  StagingSlots(const StagingSlots&) = delete;
  StagingSlots& operator=(const StagingSlots&) = delete;
  bool ok() const { return ok_ && count_ == kStagingSlotsPerCall; }
  char* slot(int i) const { return slots_[i]; }
  cudaEvent_t event(int i) const { return events_[i]; }

private:
  static std::mutex& CacheMutex()
  {
    static std::mutex mutex;
    return mutex;
  }
  static std::vector<void*>& Cache()
  {
    static std::vector<void*> cache;
    return cache;
  }
  char* slots_[kStagingSlotsPerCall] = {};
  cudaEvent_t events_[kStagingSlotsPerCall] = {};
  int count_ = 0;
  bool ok_ = true;
};

inline bool IsPageableHostPointer(const void* pointer)
{
  cudaPointerAttributes attributes{};
  if (cudaPointerGetAttributes(&attributes, pointer) != cudaSuccess)
  {
    cudaGetLastError();
    return true;
  }
  return attributes.type == cudaMemoryTypeUnregistered;
}

// ------------------------------------------------------------------------------------------------
// The two directions. Both issue their DMA on `stream`; ToDevice returns when every piece has
// been ISSUED (the host copy into the slots is done, the DMA may still run), ToHost returns when
// the data is in the caller's buffer (it synchronises on its own events only).
// Rows: `rows` rows of `row_bytes` bytes, host pitch / device pitch in bytes.
// ------------------------------------------------------------------------------------------------
class StagedTransfer
{
public:
  // Copies host -> device. Small or pinned sources go straight through cudaMemcpy2DAsync.
  cudaError_t ToDevice(char* d_dst, size_t d_pitch, const char* h_src, size_t h_pitch,
                       size_t row_bytes, size_t rows, cudaStream_t stream)
  {
    constexpr size_t kRow = size_t{1} << 20;
    const size_t total = row_bytes * rows;
    if (d_pitch == row_bytes && h_pitch == row_bytes && total > kRow)
    {
      // contiguous: 1 MiB rows (pieces then always fit a slot) plus a short tail
      const size_t main = total / kRow * kRow;
      cudaError_t status = ToDeviceRows(d_dst, kRow, h_src, kRow, kRow, main / kRow, stream);
      if (status == cudaSuccess && main < total)
      {
        status = cudaMemcpyAsync(d_dst + main, h_src + main, total - main,
                                 cudaMemcpyHostToDevice, stream);
      }
      return status;
    }
    return ToDeviceRows(d_dst, d_pitch, h_src, h_pitch, row_bytes, rows, stream);
  }

  // Copies device -> host and waits until the data is in h_dst.
  cudaError_t ToHost(char* h_dst, size_t h_pitch, const char* d_src, size_t d_pitch,
                     size_t row_bytes, size_t rows, cudaStream_t stream)
  {
    constexpr size_t kRow = size_t{1} << 20;
    const size_t total = row_bytes * rows;
    if (d_pitch == row_bytes && h_pitch == row_bytes && total > kRow)
    {
      const size_t main = total / kRow * kRow;
      if (main < total)
      {
        const cudaError_t tail = cudaMemcpyAsync(h_dst + main, d_src + main, total - main,
                                                 cudaMemcpyDeviceToHost, stream);
        if (tail != cudaSuccess)
        {
          return tail;
        }
      }
      const cudaError_t status = ToHostRows(h_dst, kRow, d_src, kRow, kRow, main / kRow, stream);
      return status != cudaSuccess ? status : cudaStreamSynchronize(stream);
    }
    return ToHostRows(h_dst, h_pitch, d_src, d_pitch, row_bytes, rows, stream);
  }

  // The host side of the staging runs on the calling thread alone instead of the shared workers.
  void UseCallingThreadOnly() { inline_copies_ = true; }

  // True when ToHost(h_dst, ...) would go through the slots (and therefore block the caller).
  bool WouldStage(const void* host, size_t bytes) { return UseStaging(host, bytes); }

private:
  cudaError_t ToDeviceRows(char* d_dst, size_t d_pitch, const char* h_src, size_t h_pitch,
                           size_t row_bytes, size_t rows, cudaStream_t stream)
  {
    if (!UseStaging(h_src, row_bytes * rows) || row_bytes > kStagingSlotBytes)
    {
      return Direct(d_dst, d_pitch, h_src, h_pitch, row_bytes, rows, cudaMemcpyHostToDevice,
                    stream);
    }
    const size_t rows_per_piece = RowsPerPiece(row_bytes);
    for (size_t row = 0; row < rows; row += rows_per_piece)
    {
      const size_t piece_rows = std::min(rows_per_piece, rows - row);
      const int s = next_slot_;
      next_slot_ = (next_slot_ + 1) % kStagingSlotsPerCall;
      if (used_[s])
      {
        const cudaError_t waited = cudaEventSynchronize(slots_.event(s));
        if (waited != cudaSuccess)
        {
          return waited;
        }
      }
      HostCopy(slots_.slot(s), row_bytes, h_src + row * h_pitch, h_pitch, row_bytes, piece_rows);
      cudaError_t status = Direct(d_dst + row * d_pitch, d_pitch, slots_.slot(s), row_bytes,
                                  row_bytes, piece_rows, cudaMemcpyHostToDevice, stream);
      if (status == cudaSuccess)
      {
        status = cudaEventRecord(slots_.event(s), stream);
      }
      if (status != cudaSuccess)
      {
        return status;
      }
      used_[s] = true;
    }
    return cudaSuccess;
  }

  cudaError_t ToHostRows(char* h_dst, size_t h_pitch, const char* d_src, size_t d_pitch,
                         size_t row_bytes, size_t rows, cudaStream_t stream)
  {
    if (!UseStaging(h_dst, row_bytes * rows) || row_bytes > kStagingSlotBytes)
    {
      const cudaError_t status = Direct(h_dst, h_pitch, d_src, d_pitch, row_bytes, rows,
                                        cudaMemcpyDeviceToHost, stream);
      return status != cudaSuccess ? status : cudaStreamSynchronize(stream);
    }
    const size_t rows_per_piece = RowsPerPiece(row_bytes);
    struct Pending
    {
      bool valid = false;
      size_t row = 0;
      size_t rows = 0;
    } pending[kStagingSlotsPerCall];
    const auto drain = [&](int s) -> cudaError_t
    {
      if (!pending[s].valid)
      {
        return cudaSuccess;
      }
      const cudaError_t waited = cudaEventSynchronize(slots_.event(s));
      if (waited != cudaSuccess)
      {
        return waited;
      }
      HostCopy(h_dst + pending[s].row * h_pitch, h_pitch, slots_.slot(s), row_bytes, row_bytes,
               pending[s].rows);
      pending[s].valid = false;
      return cudaSuccess;
    };
    int s = 0;
    for (size_t row = 0; row < rows; row += rows_per_piece)
    {
      const size_t piece_rows = std::min(rows_per_piece, rows - row);
      if (used_[s])
      {
        // (a slot last used by ToDevice: its DMA must have read it before it is overwritten)
        const cudaError_t waited = cudaEventSynchronize(slots_.event(s));
        if (waited != cudaSuccess)
        {
          return waited;
        }
        used_[s] = false;
      }
      cudaError_t status = drain(s);
      if (status == cudaSuccess)
      {
        status = Direct(slots_.slot(s), row_bytes, d_src + row * d_pitch, d_pitch, row_bytes,
                        piece_rows, cudaMemcpyDeviceToHost, stream);
      }
      if (status == cudaSuccess)
      {
        status = cudaEventRecord(slots_.event(s), stream);
      }
      if (status != cudaSuccess)
      {
        return status;
      }
      pending[s] = Pending{true, row, piece_rows};
      s = (s + 1) % kStagingSlotsPerCall;
    }
    for (int i = 0; i < kStagingSlotsPerCall; i++)
    {
      const cudaError_t status = drain((s + i) % kStagingSlotsPerCall);
      if (status != cudaSuccess)
      {
        return status;
      }
    }
    return cudaSuccess;
  }

  static constexpr size_t kMinStagedBytes = size_t{8} << 20;

  bool UseStaging(const void* host, size_t bytes)
  {
    return bytes >= kMinStagedBytes && slots_.ok() && IsPageableHostPointer(host);
  }

  static size_t RowsPerPiece(size_t row_bytes)
  {
    return std::max<size_t>(1, kStagingSlotBytes / std::max<size_t>(1, row_bytes));
  }

  static cudaError_t Direct(char* dst, size_t dst_pitch, const char* src, size_t src_pitch,
                            size_t row_bytes, size_t rows, cudaMemcpyKind kind,
                            cudaStream_t stream)
  {
    if (dst_pitch == row_bytes && src_pitch == row_bytes)
    {
      return cudaMemcpyAsync(dst, src, row_bytes * rows, kind, stream);
    }
    return cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, row_bytes, rows, kind, stream);
  }

  void HostCopy(char* dst, size_t dst_pitch, const char* src, size_t src_pitch, size_t row_bytes,
                size_t rows)
  {
    if (inline_copies_)
    {
      HostCopyPool::Copy2DInline(dst, dst_pitch, src, src_pitch, row_bytes, rows);
    }
    else
    {
      HostCopyPool::Instance().Copy2D(dst, dst_pitch, src, src_pitch, row_bytes, rows);
    }
  }

  bool inline_copies_ = false;
  StagingSlots slots_;
  bool used_[kStagingSlotsPerCall] = {};
  int next_slot_ = 0;
};
}  // namespace vgt_b200
