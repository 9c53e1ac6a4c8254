// Device code of the exact signed squared EDT (sm_100a): the z scan and the pieces shared with the
// strided-axis envelope kernel. Included by edt_kernels.cu only. See edt_kernels.cuh for the data representation.
#pragma once

#include <cstdint>

#include "edt_kernels.cuh"

namespace vgt_b200
{
namespace edt
{
namespace
{
// ================================================================================================
// Pass A: the contiguous (z) axis. Inputs are binary, so no envelope is needed: the answer is the
// distance to the nearest opposite-class voxel along the line, found with ballots + bit scans.
// One warp owns one line; every global access is a 128-byte coalesced row segment.
// Replaces, for both fields at once: the marking loop (sdfgen.hpp:57-74) and the Z-axis loop of
// ComputeDistanceFieldTransformInPlace (sdfgen.cpp:354-390).
// ================================================================================================
constexpr int kScanWarpsPerBlock = 8;
constexpr int kFar = 1 << 20;        // "no such voxel" position offset
constexpr int kFarThreshold = 1 << 19;

__device__ __forceinline__ bool IsFilled(float occupancy, int unknown_is_filled)
{
  // occupancy_map.hpp:188-193 compares the float (promoted to double) against 0.5; the float
  // comparison is the same predicate because 0.5 is exactly representable.
  return (occupancy > 0.5f) || (unknown_is_filled != 0 && occupancy == 0.5f);
}

__device__ __forceinline__ bool IsFilled(uint8_t mask, int) { return mask != 0; }

__device__ __forceinline__ float LoadStreaming(const float* p) { return __ldcs(p); }
__device__ __forceinline__ uint8_t LoadStreaming(const uint8_t* p) { return __ldcs(p); }

__device__ __forceinline__ uint32_t ValidBits(int word, int length)
{
  const int remaining = length - (word << 5);
  return (remaining >= 32) ? 0xffffffffu : ((1u << remaining) - 1u);
}

// For every 32-voxel word of a line: position of the last filled / free voxel before the word and
// of the first one after it (kFar offsets when there is none). Warp scans over 32 words at a time
// with a carry, so lines of any length work. `words` and the four tables live in shared memory.
__device__ __forceinline__ void BuildWordTables(
    const uint32_t* words, int32_t* last_filled_before, int32_t* last_free_before,
    int32_t* first_filled_after, int32_t* first_free_after, int num_words, int length, int lane)
{
  {
    int carry_filled = -kFar;
    int carry_free = -kFar;
    for (int base = 0; base < num_words; base += kWarp)
    {
      const int j = base + lane;
      uint32_t filled_bits = 0;
      uint32_t free_bits = 0;
      if (j < num_words)
      {
        const uint32_t valid = ValidBits(j, length);
        filled_bits = words[j] & valid;
        free_bits = ~words[j] & valid;
      }
      int last_filled = filled_bits ? (j << 5) + 31 - __clz(filled_bits) : -kFar;
      int last_free = free_bits ? (j << 5) + 31 - __clz(free_bits) : -kFar;
#pragma unroll
      for (int offset = 1; offset < kWarp; offset <<= 1)
      {
        const int other_filled = __shfl_up_sync(0xffffffffu, last_filled, offset);
        const int other_free = __shfl_up_sync(0xffffffffu, last_free, offset);
        if (lane >= offset)
        {
          last_filled = max(last_filled, other_filled);
          last_free = max(last_free, other_free);
        }
      }
      int before_filled = __shfl_up_sync(0xffffffffu, last_filled, 1);
      int before_free = __shfl_up_sync(0xffffffffu, last_free, 1);
      if (lane == 0)
      {
        before_filled = -kFar;
        before_free = -kFar;
      }
      before_filled = max(before_filled, carry_filled);
      before_free = max(before_free, carry_free);
      if (j < num_words)
      {
        last_filled_before[j] = before_filled;
        last_free_before[j] = before_free;
      }
      carry_filled = max(carry_filled, __shfl_sync(0xffffffffu, last_filled, 31));
      carry_free = max(carry_free, __shfl_sync(0xffffffffu, last_free, 31));
    }
  }
  {
    int carry_filled = kFar;
    int carry_free = kFar;
    const int num_groups = (num_words + kWarp - 1) / kWarp;
    for (int group = num_groups - 1; group >= 0; group--)
    {
      const int j = group * kWarp + lane;
      uint32_t filled_bits = 0;
      uint32_t free_bits = 0;
      if (j < num_words)
      {
        const uint32_t valid = ValidBits(j, length);
        filled_bits = words[j] & valid;
        free_bits = ~words[j] & valid;
      }
      int first_filled = filled_bits ? (j << 5) + __ffs(filled_bits) - 1 : kFar;
      int first_free = free_bits ? (j << 5) + __ffs(free_bits) - 1 : kFar;
#pragma unroll
      for (int offset = 1; offset < kWarp; offset <<= 1)
      {
        const int other_filled = __shfl_down_sync(0xffffffffu, first_filled, offset);
        const int other_free = __shfl_down_sync(0xffffffffu, first_free, offset);
        if (lane + offset < kWarp)
        {
          first_filled = min(first_filled, other_filled);
          first_free = min(first_free, other_free);
        }
      }
      int after_filled = __shfl_down_sync(0xffffffffu, first_filled, 1);
      int after_free = __shfl_down_sync(0xffffffffu, first_free, 1);
      if (lane == kWarp - 1)
      {
        after_filled = kFar;
        after_free = kFar;
      }
      after_filled = min(after_filled, carry_filled);
      after_free = min(after_free, carry_free);
      if (j < num_words)
      {
        first_filled_after[j] = after_filled;
        first_free_after[j] = after_free;
      }
      carry_filled = min(carry_filled, __shfl_sync(0xffffffffu, first_filled, 0));
      carry_free = min(carry_free, __shfl_sync(0xffffffffu, first_free, 0));
    }
  }
}

template <typename In>
__global__ void __launch_bounds__(kScanWarpsPerBlock * kWarp) ScanContiguousAxisKernel(
    const In* __restrict__ in, uint32_t* __restrict__ out, int64_t num_lines, int32_t length,
    int unknown_is_filled)
{
  extern __shared__ uint32_t scan_smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_words = (length + 31) >> 5;
  const int64_t line = static_cast<int64_t>(blockIdx.x) * kScanWarpsPerBlock + warp;
  if (line >= num_lines)
  {
    return;  // warp-uniform
  }
  uint32_t* words = scan_smem + warp * 5 * num_words;
  int32_t* last_filled_before = reinterpret_cast<int32_t*>(words + num_words);
  int32_t* last_free_before = last_filled_before + num_words;
  int32_t* first_filled_after = last_free_before + num_words;
  int32_t* first_free_after = first_filled_after + num_words;

  const In* src = in + line * length;
  uint32_t* dst = out + line * length;

  // 1. classify: one ballot word per 32 voxels.
#pragma unroll 4
  for (int w = 0; w < num_words; w++)
  {
    const int z = (w << 5) + lane;
    bool filled = false;
    if (z < length)
    {
      filled = IsFilled(LoadStreaming(src + z), unknown_is_filled);
    }
    const uint32_t word = __ballot_sync(0xffffffffu, filled);
    if (lane == 0)
    {
      words[w] = word;
    }
  }
  __syncwarp();

  // 2. per-word tables of the nearest filled / free voxel outside the word.
  BuildWordTables(words, last_filled_before, last_free_before, first_filled_after,
                  first_free_after, num_words, length, lane);
  __syncwarp();

  // 3. per voxel: nearest opposite-class voxel inside the word (bit scan) or outside (tables).
  for (int w = 0; w < num_words; w++)
  {
    const int z = (w << 5) + lane;
    const uint32_t word = words[w];
    const uint32_t valid = ValidBits(w, length);
    const uint32_t filled = (word >> lane) & 1u;
    const uint32_t opposite = (filled ? ~word : word) & valid;
    const uint32_t below = opposite & ((1u << lane) - 1u);
    const uint32_t above = (lane == 31) ? 0u : (opposite >> (lane + 1));
    const int left = below ? (lane - (31 - __clz(below)))
                           : (z - (filled ? last_free_before[w] : last_filled_before[w]));
    const int right = above ? __ffs(above)
                            : ((filled ? first_free_after[w] : first_filled_after[w]) - z);
    const int nearest = min(left, right);
    const uint32_t squared =
        (nearest >= kFarThreshold) ? kNone : static_cast<uint32_t>(nearest * nearest);
    if (z < length)
    {
      dst[z] = (filled << 31) | squared;
    }
  }
}

// Four voxels per lane: 128-bit loads and stores, 128 voxels per warp iteration. Needs lines that
// start 16-byte aligned (length % 4 == 0). Same algorithm as above; a lane's four class bits are
// merged into its group's 32-bit word with one redux.or over the 8 lanes that share the word.
struct Float4Source
{
  using Vector = float4;
  __device__ static __forceinline__ uint32_t Nibble(const float4& v, int unknown_is_filled)
  {
    return (IsFilled(v.x, unknown_is_filled) ? 1u : 0u) | (IsFilled(v.y, unknown_is_filled) ? 2u : 0u)
        | (IsFilled(v.z, unknown_is_filled) ? 4u : 0u) | (IsFilled(v.w, unknown_is_filled) ? 8u : 0u);
  }
};

struct Uchar4Source
{
  using Vector = uchar4;
  __device__ static __forceinline__ uint32_t Nibble(const uchar4& v, int)
  {
    return (v.x ? 1u : 0u) | (v.y ? 2u : 0u) | (v.z ? 4u : 0u) | (v.w ? 8u : 0u);
  }
};

template <typename Source>
__global__ void __launch_bounds__(kScanWarpsPerBlock * kWarp) ScanContiguousAxisVec4Kernel(
    const typename Source::Vector* __restrict__ in, uint4* __restrict__ out, int64_t num_lines,
    int32_t length, int unknown_is_filled)
{
  using Vector = typename Source::Vector;
  extern __shared__ uint32_t scan_smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_words = (length + 31) >> 5;
  const int64_t line = static_cast<int64_t>(blockIdx.x) * kScanWarpsPerBlock + warp;
  if (line >= num_lines)
  {
    return;  // warp-uniform
  }
  uint32_t* words = scan_smem + warp * 5 * num_words;
  int32_t* last_filled_before = reinterpret_cast<int32_t*>(words + num_words);
  int32_t* last_free_before = last_filled_before + num_words;
  int32_t* first_filled_after = last_free_before + num_words;
  int32_t* first_free_after = first_filled_after + num_words;

  const int vectors = length >> 2;  // per line
  const Vector* src = in + line * vectors;
  uint4* dst = out + line * vectors;
  const int group = lane >> 3;                       // which of the 4 words of this iteration
  const int bit0 = (lane & 7) << 2;                  // first of this lane's 4 bits in that word
  const unsigned group_mask = 0xffu << (group << 3);
  const int iterations = (length + 127) >> 7;

  // 1. classify: 4 voxels per lane, one 32-bit class word per 8 lanes.
#pragma unroll 2
  for (int it = 0; it < iterations; it++)
  {
    const int vector_index = (it << 5) + lane;
    uint32_t nibble = 0;
    if (vector_index < vectors)
    {
      nibble = Source::Nibble(__ldcs(src + vector_index), unknown_is_filled);
    }
    const uint32_t word = __reduce_or_sync(group_mask, nibble << bit0);
    const int w = (it << 2) + group;
    if ((lane & 7) == 0 && w < num_words)
    {
      words[w] = word;
    }
  }
  __syncwarp();

  // 2. per-word tables of the nearest filled / free voxel outside the word.
  BuildWordTables(words, last_filled_before, last_free_before, first_filled_after,
                  first_free_after, num_words, length, lane);
  __syncwarp();

  // 3. per voxel: nearest opposite-class voxel inside the word (bit scan) or outside (tables).
  for (int it = 0; it < iterations; it++)
  {
    const int vector_index = (it << 5) + lane;
    const int w = (it << 2) + group;
    if (vector_index >= vectors)
    {
      continue;
    }
    const uint32_t word = words[w];
    const uint32_t valid = ValidBits(w, length);
    const int before_filled = last_filled_before[w];
    const int before_free = last_free_before[w];
    const int after_filled = first_filled_after[w];
    const int after_free = first_free_after[w];
    uint32_t results[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
      const int bit = bit0 + k;
      const int z = (w << 5) + bit;
      const uint32_t filled = (word >> bit) & 1u;
      const uint32_t opposite = (filled ? ~word : word) & valid;
      const uint32_t below = opposite & ((1u << bit) - 1u);
      const uint32_t above = (bit == 31) ? 0u : (opposite >> (bit + 1));
      const int left = below ? (bit - (31 - __clz(below))) : (z - (filled ? before_free : before_filled));
      const int right = above ? __ffs(above) : ((filled ? after_free : after_filled) - z);
      const int nearest = min(left, right);
      const uint32_t squared =
          (nearest >= kFarThreshold) ? kNone : static_cast<uint32_t>(nearest * nearest);
      results[k] = (filled << 31) | squared;
    }
    dst[vector_index] = make_uint4(results[0], results[1], results[2], results[3]);
  }
}

// ================================================================================================
// Shared pieces of the strided-axis envelope kernel (edt_envelope_inplace.cuh): the site type, the
// exact pop test, the finalize expression and the order-preserving min/max keys.
// ================================================================================================
struct Site
{
  int32_t v;  // position on the line
  int32_t h;  // f(v) + v*v
};

constexpr int32_t kNoSitePosition = 0x7fffffff;
constexpr int32_t kNoSiteHeight = 0x3fffffff;

// True when the middle site never owns a point of the lower envelope given its neighbours:
// crossing(below, middle) >= crossing(middle, above), cross-multiplied (all denominators > 0).
// This is the F-H pop test "s <= z[k]" (sdfgen.cpp:190-194) in exact integer arithmetic.
__device__ __forceinline__ bool MiddleIsHidden(const Site& below, const Site& middle,
                                               const Site& above)
{
  const long long lhs =
      static_cast<long long>(middle.h - below.h) * static_cast<long long>(above.v - middle.v);
  const long long rhs =
      static_cast<long long>(above.h - middle.h) * static_cast<long long>(middle.v - below.v);
  return lhs >= rhs;
}

__device__ __forceinline__ uint32_t OrderedKey(float value)
{
  const uint32_t bits = __float_as_uint(value);
  return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
}

__device__ __forceinline__ unsigned long long OrderedKey(double value)
{
  const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(value));
  return (bits & 0x8000000000000000ull) ? ~bits : (bits | 0x8000000000000000ull);
}

template <typename T>
__device__ __forceinline__ T PositiveInfinity();
template <>
__device__ __forceinline__ float PositiveInfinity<float>()
{
  return __int_as_float(0x7f800000);
}
template <>
__device__ __forceinline__ double PositiveInfinity<double>()
{
  return __longlong_as_double(0x7ff0000000000000ll);
}
template <>
__device__ __forceinline__ uint32_t PositiveInfinity<uint32_t>()
{
  return 0u;  // unused: packed mode has no min/max
}

template <typename Out>
__device__ __forceinline__ Out SignedDistanceOf(uint32_t filled, uint32_t squared,
                                                double resolution)
{
  // sdfgen.hpp:98-105: sqrt(sq) * resolution in double, cast, sign by class. Exactly one of the
  // reference's two fields is zero at any voxel, so the difference is +/- this value.
  Out magnitude;
  if (squared == kNone)
  {
    magnitude = PositiveInfinity<Out>();
  }
  else
  {
    magnitude = static_cast<Out>(__dmul_rn(__dsqrt_rn(static_cast<double>(squared)), resolution));
  }
  return filled ? -magnitude : magnitude;
}

// The same value through the per-call magnitude table when the squared distance is inside it.
template <typename Out>
__device__ __forceinline__ Out SignedDistanceFromTable(uint32_t filled, uint32_t squared,
                                                       double resolution, const Out* table,
                                                       uint32_t table_size)
{
  if (squared < table_size)
  {
    const Out magnitude = __ldg(table + squared);
    return filled ? -magnitude : magnitude;
  }
  return SignedDistanceOf<Out>(filled, squared, resolution);
}

template <typename Out>
__global__ void BuildMagnitudeTableKernel(Out* table, uint32_t size, double resolution)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < size)
  {
    table[s] = SignedDistanceOf<Out>(0u, s, resolution);
  }
}

enum EnvelopeMode
{
  kEmitPacked = 0,  // int32 sign-fused word out
  kEmitFloat = 1,   // finalize to float SDF
  kEmitDouble = 2   // finalize to double SDF
};

template <int kMode>
struct OutputOf
{
  using Type = uint32_t;
  using Key = uint32_t;
};
template <>
struct OutputOf<kEmitFloat>
{
  using Type = float;
  using Key = uint32_t;
};
template <>
struct OutputOf<kEmitDouble>
{
  using Type = double;
  using Key = unsigned long long;
};

// Decodes the ordered keys back into values (one thread).
template <typename Out, typename Key>
__global__ void DecodeMinMaxKernel(const Key* keys, Out* min_max);

template <>
__global__ void DecodeMinMaxKernel<float, uint32_t>(const uint32_t* keys, float* min_max)
{
  for (int i = 0; i < 2; i++)
  {
    const uint32_t key = keys[i];
    const uint32_t bits = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    min_max[i] = __uint_as_float(bits);
  }
}

template <>
__global__ void DecodeMinMaxKernel<double, unsigned long long>(
    const unsigned long long* keys, double* min_max)
{
  for (int i = 0; i < 2; i++)
  {
    const unsigned long long key = keys[i];
    const unsigned long long bits =
        (key & 0x8000000000000000ull) ? (key & 0x7fffffffffffffffull) : ~key;
    min_max[i] = __longlong_as_double(static_cast<long long>(bits));
  }
}

template <typename Key>
__global__ void ResetMinMaxKeysKernel(Key* keys)
{
  keys[0] = ~static_cast<Key>(0);
  keys[1] = 0;
}

// Splits the sign-fused word into the reference's two squared fields (parity hook only).
__global__ void SplitFieldsKernel(const uint32_t* packed, int64_t count, int32_t* to_filled,
                                  int32_t* to_free)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count)
  {
    return;
  }
  const uint32_t word = packed[i];
  const int32_t value = static_cast<int32_t>(word & kNone);  // kNone == INT32_MAX == SQ_INF
  const bool filled = (word >> 31) != 0;
  to_filled[i] = filled ? 0 : value;
  to_free[i] = filled ? value : 0;
}

}  // namespace
}  // namespace edt
}  // namespace vgt_b200
