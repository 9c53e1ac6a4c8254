// TEST INFRASTRUCTURE ONLY -- stand-in for common_robotics_utilities/parallelism.hpp (the real
// library is an unvendored, unpinned dependency of the reference: package.xml.ros2:12).
// Only what signed_distance_field_generation.cpp and cpu_pointcloud_voxelization.cpp use:
// DegreeOfParallelism, ThreadWorkRange, StaticParallelForRangeLoop (a static split of
// [start, end) into one range per thread) and StaticParallelForIndexLoop (the same split, the
// functor called per index).
#pragma once

#include <cstdint>
#include <stdexcept>

#include <common_robotics_utilities/openmp_helpers.hpp>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace common_robotics_utilities
{
namespace parallelism
{
class DegreeOfParallelism
{
public:
  static DegreeOfParallelism None() { return DegreeOfParallelism(1); }
  static DegreeOfParallelism FromOmp()
  {
#ifdef _OPENMP
    return DegreeOfParallelism(omp_get_max_threads());
#else
    return DegreeOfParallelism(1);
#endif
  }
  DegreeOfParallelism() : DegreeOfParallelism(1) {}
  explicit DegreeOfParallelism(int32_t num_threads) : num_threads_(num_threads)
  {
    if (num_threads_ < 1) { throw std::invalid_argument("num_threads must be >= 1"); }
  }
  bool IsParallel() const { return num_threads_ > 1; }
  int32_t GetNumThreads() const { return num_threads_; }

private:
  int32_t num_threads_ = 1;
};

enum class ParallelForBackend { BEST_AVAILABLE, OPENMP, ASYNC };

class ThreadWorkRange
{
public:
  ThreadWorkRange(int64_t range_start, int64_t range_end, int32_t thread_num)
      : range_start_(range_start), range_end_(range_end), thread_num_(thread_num) {}
  int64_t GetRangeStart() const { return range_start_; }
  int64_t GetRangeEnd() const { return range_end_; }
  int32_t GetThreadNum() const { return thread_num_; }

private:
  int64_t range_start_;
  int64_t range_end_;
  int32_t thread_num_;
};

template <typename Functor>
void StaticParallelForRangeLoop(
    const DegreeOfParallelism& parallelism, int64_t range_start, int64_t range_end,
    const Functor& functor, ParallelForBackend = ParallelForBackend::BEST_AVAILABLE)
{
  const int32_t threads = parallelism.GetNumThreads();
  const int64_t total = range_end - range_start;
  const int64_t base = total / threads;
  const int64_t extra = total % threads;
  if (threads == 1)
  {
    // no parallel region: an exception thrown by the functor reaches the caller (the mesh
    // rasterizer throws from inside its loop)
    functor(ThreadWorkRange(range_start, range_end, 0));
    return;
  }
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static)
#endif
  for (int32_t t = 0; t < threads; t++)
  {
    const int64_t begin = range_start + t * base + (t < extra ? t : extra);
    const int64_t end = begin + base + (t < extra ? 1 : 0);
    functor(ThreadWorkRange(begin, end, t));
  }
}
// (device_pointcloud_voxelization.cpp:147-149: one item per cloud; the order in which items are
// handed out does not matter to the caller, a static split serves)
template <typename Functor>
void DynamicParallelForIndexLoop(
    const DegreeOfParallelism& parallelism, int64_t range_start, int64_t range_end,
    const Functor& functor, ParallelForBackend backend = ParallelForBackend::BEST_AVAILABLE);

template <typename Functor>
void StaticParallelForIndexLoop(
    const DegreeOfParallelism& parallelism, int64_t range_start, int64_t range_end,
    const Functor& functor, ParallelForBackend backend = ParallelForBackend::BEST_AVAILABLE)
{
  StaticParallelForRangeLoop(
      parallelism, range_start, range_end,
      [&](const ThreadWorkRange& range)
      {
        for (int64_t index = range.GetRangeStart(); index < range.GetRangeEnd(); index++)
        {
          functor(range.GetThreadNum(), index);
        }
      },
      backend);
}
template <typename Functor>
void DynamicParallelForIndexLoop(
    const DegreeOfParallelism& parallelism, int64_t range_start, int64_t range_end,
    const Functor& functor, ParallelForBackend backend)
{
  StaticParallelForIndexLoop(parallelism, range_start, range_end, functor, backend);
}
}  // namespace parallelism
}  // namespace common_robotics_utilities
