// Device code of the exact signed squared EDT (sm_100a): the z scan and the strided-axis envelope
// kernels. Included by edt_kernels.cu only. See edt_kernels.cuh for the data representation.
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

  // 2. for every word: position of the last filled / free voxel before it and the first after it
  //    (warp scans over 32 words at a time with a carry).
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

// ================================================================================================
// Passes B and C: a strided axis. One lane owns one line, a warp owns kWarp (or fewer, for very
// long lines) adjacent lines, so every global access of the warp is one contiguous row segment.
// Each lane runs the run-decomposed Felzenszwalb-Huttenlocher envelope with its stack in shared
// memory ([slot][lane] layout: lanes never collide on a bank), then sweeps the line once more to
// emit results. Replaces the X / Y loops of ComputeDistanceFieldTransformInPlace
// (sdfgen.cpp:276-351) and the 1-D transforms (sdfgen.cpp:85-226) for both fields at once; in
// finalize mode also the combine loop (sdfgen.hpp:85-108) and Lock()'s min/max (sdf.hpp:765-787).
// ================================================================================================
struct Site
{
  int32_t v;  // position on the line
  int32_t h;  // f(v) + v*v
};

constexpr int32_t kNoSitePosition = 0x7fffffff;
constexpr int32_t kNoSiteHeight = 0x3fffffff;

template <int kEntryBytes>
struct EntryCodec;

// Packed entry for lines up to 1024 voxels whose finite inputs stay below 2^22.
template <>
struct EntryCodec<4>
{
  using Storage = uint32_t;
  static constexpr int kPositionBits = 10;
  __device__ static __forceinline__ Storage Pack(int32_t v, int32_t f, int32_t)
  {
    return (static_cast<uint32_t>(f) << kPositionBits) | static_cast<uint32_t>(v);
  }
  __device__ static __forceinline__ Site Unpack(Storage e)
  {
    const int32_t v = static_cast<int32_t>(e & ((1u << kPositionBits) - 1u));
    const int32_t f = static_cast<int32_t>(e >> kPositionBits);
    return Site{v, f + v * v};
  }
};

template <>
struct EntryCodec<8>
{
  using Storage = int2;
  __device__ static __forceinline__ Storage Pack(int32_t v, int32_t, int32_t h)
  {
    return make_int2(v, h);
  }
  __device__ static __forceinline__ Site Unpack(Storage e) { return Site{e.x, e.y}; }
};

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

constexpr int kPrefetch = 8;

template <int kEntryBytes, int kMode>
__global__ void __launch_bounds__(kWarp) EnvelopeAxisKernel(
    const uint32_t* in, typename OutputOf<kMode>::Type* out, LineFamily family,
    int lanes_per_tile, FinalizeParams finalize, typename OutputOf<kMode>::Key* min_max_keys)
{
  using Codec = EntryCodec<kEntryBytes>;
  using Entry = typename Codec::Storage;
  using Out = typename OutputOf<kMode>::Type;
  extern __shared__ __align__(16) unsigned char envelope_smem[];

  const int lane = threadIdx.x;
  const int length = family.length;
  const int num_words = (length + 31) >> 5;
  Entry* stack = reinterpret_cast<Entry*>(envelope_smem);                  // [length][lanes]
  uint32_t* class_words = reinterpret_cast<uint32_t*>(stack + static_cast<size_t>(length) * lanes_per_tile);

  const int64_t tiles_per_outer = (family.inner_count + lanes_per_tile - 1) / lanes_per_tile;
  const int64_t outer = blockIdx.x / tiles_per_outer;
  const int64_t tile = blockIdx.x - outer * tiles_per_outer;
  const int64_t column = tile * lanes_per_tile + lane;
  const bool active = (lane < lanes_per_tile) && (column < family.inner_count);
  const int64_t first = outer * family.outer_stride + column;
  const int64_t stride = family.line_stride;

  Out lane_min = PositiveInfinity<Out>();
  Out lane_max = -PositiveInfinity<Out>();

  if (active)
  {
    const uint32_t* src = in + first;

    // ------------------------------------------------------------------ phase 1: build stacks
    int slot = 0;        // next free stack slot (runs are stored back to back)
    int depth = 0;       // stored sites of the current run
    bool has_left = false;
    Site left_zero{0, 0};  // zero-height site just before the current run
    Site top{0, 0};
    Site below{0, 0};
    uint32_t previous_class = 0;
    uint32_t word_accumulator = 0;

    const auto pop_hidden = [&](const Site& incoming)
    {
      while ((depth >= 2 || (depth == 1 && has_left)) && MiddleIsHidden(below, top, incoming))
      {
        depth--;
        slot--;
        top = below;
        if (depth >= 2)
        {
          below = Codec::Unpack(stack[static_cast<size_t>(slot - 2) * lanes_per_tile + lane]);
        }
        else
        {
          below = left_zero;  // only meaningful when depth == 1 && has_left
        }
      }
    };

    uint32_t current[kPrefetch];
    uint32_t upcoming[kPrefetch];
#pragma unroll
    for (int u = 0; u < kPrefetch; u++)
    {
      current[u] = (u < length) ? __ldcg(src + static_cast<int64_t>(u) * stride) : 0u;
    }
    for (int q0 = 0; q0 < length; q0 += kPrefetch)
    {
#pragma unroll
      for (int u = 0; u < kPrefetch; u++)
      {
        const int q = q0 + kPrefetch + u;
        upcoming[u] = (q < length) ? __ldcg(src + static_cast<int64_t>(q) * stride) : 0u;
      }
#pragma unroll
      for (int u = 0; u < kPrefetch; u++)
      {
        const int q = q0 + u;
        if (q < length)
        {
          const uint32_t word = current[u];
          const uint32_t filled = word >> 31;
          const uint32_t value = word & kNone;
          word_accumulator |= filled << (q & 31);
          if ((q & 31) == 31 || q == length - 1)
          {
            class_words[static_cast<size_t>(q >> 5) * lanes_per_tile + lane] = word_accumulator;
            word_accumulator = 0;
          }
          if (q > 0 && filled != previous_class)
          {
            // The run ends: voxel q is a zero-height site for it. It hides what it hides, but is
            // not stored (phase 2 re-creates it from the class bits).
            pop_hidden(Site{q, q * q});
            depth = 0;
            has_left = true;
            left_zero = Site{q - 1, (q - 1) * (q - 1)};
            top = left_zero;
          }
          previous_class = filled;
          if (value != kNone)
          {
            const Site incoming{q, static_cast<int32_t>(value) + q * q};
            pop_hidden(incoming);
            stack[static_cast<size_t>(slot) * lanes_per_tile + lane] =
                Codec::Pack(q, static_cast<int32_t>(value), incoming.h);
            below = top;
            top = incoming;
            depth++;
            slot++;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kPrefetch; u++)
      {
        current[u] = upcoming[u];
      }
    }

    // ------------------------------------------------------------------ phase 2: sweep
    const int stored_total = slot;
    int cursor = 0;  // index of the stored site held in `pending`
    const auto load_stored = [&](int index)
    {
      if (index < stored_total)
      {
        return Codec::Unpack(stack[static_cast<size_t>(index) * lanes_per_tile + lane]);
      }
      return Site{kNoSitePosition, kNoSiteHeight};
    };
    Site pending = load_stored(0);
    Site winner{0, kNoSiteHeight};
    int run_end = 0;            // first position after the current run
    bool right_zero_used = true;
    uint32_t class_word = 0;
    previous_class = 0;

    int32_t border_yz = 0x7fffffff;
    if (kMode != kEmitPacked && finalize.add_virtual_border != 0)
    {
      const int32_t y = finalize.y_offset + static_cast<int32_t>(column / finalize.nz);
      const int32_t z = static_cast<int32_t>(column % finalize.nz);
      if (finalize.ny_total > 1)
      {
        border_yz = min(border_yz, min(y + 1, finalize.ny_total - y));
      }
      if (finalize.nz_total > 1)
      {
        border_yz = min(border_yz, min(z + 1, finalize.nz_total - z));
      }
    }

    Out* dst = out + first;
    for (int q = 0; q < length; q++)
    {
      if ((q & 31) == 0)
      {
        class_word = class_words[static_cast<size_t>(q >> 5) * lanes_per_tile + lane];
      }
      const uint32_t filled = (class_word >> (q & 31)) & 1u;
      if (q == 0 || filled != previous_class)
      {
        // A run starts at q: find where it ends from the class bits.
        int w = q >> 5;
        uint32_t different = (filled ? ~class_word : class_word) & (0xffffffffu << (q & 31));
        while (different == 0 && ++w < num_words)
        {
          const uint32_t bits = class_words[static_cast<size_t>(w) * lanes_per_tile + lane];
          different = filled ? ~bits : bits;
        }
        run_end = different ? min(length, (w << 5) + __ffs(different) - 1) : length;
        // Drop stored sites of earlier runs that the sweep never reached.
        while (pending.v < q)
        {
          cursor++;
          pending = load_stored(cursor);
        }
        right_zero_used = (run_end >= length);
        if (q > 0)
        {
          winner = Site{q - 1, (q - 1) * (q - 1)};
        }
        else if (pending.v < run_end)
        {
          winner = pending;
          cursor++;
          pending = load_stored(cursor);
        }
        else if (!right_zero_used)
        {
          winner = Site{run_end, run_end * run_end};
          right_zero_used = true;
        }
        else
        {
          winner = Site{0, kNoSiteHeight};
        }
      }
      previous_class = filled;

      // Advance while the next candidate is strictly lower at q (F-H "while z[k+1] < q").
      while (true)
      {
        Site candidate;
        bool from_stack = false;
        if (pending.v < run_end)
        {
          candidate = pending;
          from_stack = true;
        }
        else if (!right_zero_used)
        {
          candidate = Site{run_end, run_end * run_end};
        }
        else
        {
          break;
        }
        const int32_t candidate_value = candidate.h - 2 * candidate.v * q;
        const int32_t winner_value = winner.h - 2 * winner.v * q;
        if (winner.h != kNoSiteHeight && !(candidate_value < winner_value))
        {
          break;
        }
        winner = candidate;
        if (from_stack)
        {
          cursor++;
          pending = load_stored(cursor);
        }
        else
        {
          right_zero_used = true;
        }
      }

      uint32_t squared = kNone;
      if (winner.h != kNoSiteHeight)
      {
        squared = static_cast<uint32_t>(winner.h - 2 * winner.v * q + q * q);
      }

      if constexpr (kMode == kEmitPacked)
      {
        reinterpret_cast<uint32_t*>(dst)[static_cast<int64_t>(q) * stride] =
            (filled << 31) | squared;
      }
      else
      {
        if (finalize.add_virtual_border)
        {
          int32_t border = border_yz;
          if (finalize.nx_total > 1)
          {
            border = min(border, min(q + 1, finalize.nx_total - q));
          }
          if (border != 0x7fffffff)
          {
            squared = min(squared, static_cast<uint32_t>(border * border));
          }
        }
        const Out value = SignedDistanceOf<Out>(filled, squared, finalize.resolution);
        dst[static_cast<int64_t>(q) * stride] = value;
        lane_min = (value < lane_min) ? value : lane_min;
        lane_max = (value > lane_max) ? value : lane_max;
      }
    }
  }

  if constexpr (kMode != kEmitPacked)
  {
    if (min_max_keys == nullptr)
    {
      return;
    }
    using Key = typename OutputOf<kMode>::Key;
    Key key_min = OrderedKey(lane_min);
    Key key_max = OrderedKey(lane_max);
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1)
    {
      const Key other_min = __shfl_xor_sync(0xffffffffu, key_min, offset);
      const Key other_max = __shfl_xor_sync(0xffffffffu, key_max, offset);
      key_min = (other_min < key_min) ? other_min : key_min;
      key_max = (other_max > key_max) ? other_max : key_max;
    }
    if (lane == 0)
    {
      atomicMin(min_max_keys + 0, key_min);
      atomicMax(min_max_keys + 1, key_max);
    }
  }
}

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
