// sm_100a kernels + C-ABI entry points for the occupancy -> SignedDistanceField path.
// See edt_kernels.cuh for the data representation and DESIGN.md for the roofline of each kernel.
#include "edt_kernels.cuh"

#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "common.cuh"

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

// ================================================================================================
// Host-side launch helpers
// ================================================================================================
constexpr size_t kMaxDynamicSmem = 227 * 1024;

template <typename In>
int LaunchScan(const In* d_in, uint32_t* d_out, int64_t num_lines, int32_t length,
               int unknown_is_filled, cudaStream_t stream)
{
  const int num_words = (length + 31) >> 5;
  const size_t smem = sizeof(uint32_t) * 5 * num_words * kScanWarpsPerBlock;
  const int64_t blocks = (num_lines + kScanWarpsPerBlock - 1) / kScanWarpsPerBlock;
  ScanContiguousAxisKernel<In><<<static_cast<unsigned>(blocks), kScanWarpsPerBlock * kWarp, smem,
                                 stream>>>(d_in, d_out, num_lines, length, unknown_is_filled);
  VGT_CUDA_TRY(cudaGetLastError(), "ScanContiguousAxisKernel launch");
  return VGT_B200_OK;
}

template <int kEntryBytes, int kMode>
int LaunchEnvelopeVariant(const uint32_t* d_in, typename OutputOf<kMode>::Type* d_out,
                          const LineFamily& family, const FinalizeParams& finalize,
                          typename OutputOf<kMode>::Key* d_keys, cudaStream_t stream)
{
  const int num_words = (family.length + 31) >> 5;
  int lanes = kWarp;
  size_t smem = 0;
  while (true)
  {
    smem = (static_cast<size_t>(family.length) * kEntryBytes + static_cast<size_t>(num_words) * 4)
        * lanes;
    if (smem <= kMaxDynamicSmem || lanes == 1)
    {
      break;
    }
    lanes >>= 1;
  }
  if (smem > kMaxDynamicSmem)
  {
    return FailInvalid("axis of %d voxels does not fit the shared-memory envelope stack",
                       family.length);
  }
  auto kernel = EnvelopeAxisKernel<kEntryBytes, kMode>;
  VGT_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(kMaxDynamicSmem)),
               "EnvelopeAxisKernel smem attribute");
  const int64_t tiles_per_outer = (family.inner_count + lanes - 1) / lanes;
  const int64_t blocks = tiles_per_outer * family.num_outer;
  if (blocks > 0x7fffffffLL)
  {
    return FailInvalid("grid too large for one launch");
  }
  kernel<<<static_cast<unsigned>(blocks), kWarp, smem, stream>>>(d_in, d_out, family, lanes,
                                                                 finalize, d_keys);
  VGT_CUDA_TRY(cudaGetLastError(), "EnvelopeAxisKernel launch");
  return VGT_B200_OK;
}

// max_input: largest finite partial squared distance the pass can see.
template <int kMode>
int LaunchEnvelope(const uint32_t* d_in, typename OutputOf<kMode>::Type* d_out,
                   const LineFamily& family, int64_t max_input, const FinalizeParams& finalize,
                   typename OutputOf<kMode>::Key* d_keys, cudaStream_t stream)
{
  const bool packed_fits = family.length <= (1 << EntryCodec<4>::kPositionBits)
      && max_input < (int64_t{1} << (32 - EntryCodec<4>::kPositionBits));
  if (packed_fits)
  {
    return LaunchEnvelopeVariant<4, kMode>(d_in, d_out, family, finalize, d_keys, stream);
  }
  return LaunchEnvelopeVariant<8, kMode>(d_in, d_out, family, finalize, d_keys, stream);
}

inline int64_t Square(int64_t v) { return v * v; }

// Passes along z then y on an [nx, ny, nz] block (both independent per x).
template <typename In>
int RunLocalPasses(const In* d_in, int64_t nx, int64_t ny, int64_t nz, int unknown_is_filled,
                   uint32_t* d_packed, cudaStream_t stream)
{
  int status = LaunchScan<In>(d_in, d_packed, nx * ny, static_cast<int32_t>(nz),
                              unknown_is_filled, stream);
  if (status != VGT_B200_OK)
  {
    return status;
  }
  if (ny > 1)
  {
    const LineFamily along_y{nx, ny * nz, nz, nz, static_cast<int32_t>(ny)};
    status = LaunchEnvelope<kEmitPacked>(d_packed, d_packed, along_y, Square(nz - 1),
                                         FinalizeParams{}, nullptr, stream);
  }
  return status;
}

// Pass along x on an [nx, ny_local, nz] block + finalize. `d_out` may alias `d_packed` when the
// output element is 4 bytes.
template <int kMode>
int RunFinalPass(const uint32_t* d_packed, int64_t nx, int64_t ny_local, int64_t nz,
                 int64_t y_offset, int64_t ny_total, double resolution, int add_virtual_border,
                 typename OutputOf<kMode>::Type* d_out, typename OutputOf<kMode>::Type* d_min_max,
                 typename OutputOf<kMode>::Key* d_keys, cudaStream_t stream)
{
  using Out = typename OutputOf<kMode>::Type;
  using Key = typename OutputOf<kMode>::Key;
  if (d_min_max != nullptr)
  {
    ResetMinMaxKeysKernel<Key><<<1, 1, 0, stream>>>(d_keys);
  }
  const LineFamily along_x{1, 0, ny_local * nz, ny_local * nz, static_cast<int32_t>(nx)};
  FinalizeParams finalize{};
  finalize.resolution = resolution;
  finalize.add_virtual_border = add_virtual_border;
  finalize.nx_total = static_cast<int32_t>(nx);
  finalize.ny_total = static_cast<int32_t>(ny_total);
  finalize.nz_total = static_cast<int32_t>(nz);
  finalize.y_offset = static_cast<int32_t>(y_offset);
  finalize.nz = static_cast<int32_t>(nz);
  const int status = LaunchEnvelope<kMode>(
      d_packed, d_out, along_x, Square(nz - 1) + Square(ny_total - 1), finalize,
      (d_min_max != nullptr) ? d_keys : nullptr, stream);
  if (status != VGT_B200_OK)
  {
    return status;
  }
  if (d_min_max != nullptr)
  {
    DecodeMinMaxKernel<Out, Key><<<1, 1, 0, stream>>>(d_keys, d_min_max);
    VGT_CUDA_TRY(cudaGetLastError(), "DecodeMinMaxKernel launch");
  }
  return VGT_B200_OK;
}

int CheckSdfArguments(const void* in, const void* out, int64_t nx, int64_t ny, int64_t nz,
                      double resolution)
{
  if (in == nullptr || out == nullptr)
  {
    return FailInvalid("null grid pointer");
  }
  if (!ValidDims(nx, ny, nz))
  {
    return FailInvalid("grid dimensions %lld x %lld x %lld out of range [1, %d]",
                       static_cast<long long>(nx), static_cast<long long>(ny),
                       static_cast<long long>(nz), VGT_B200_MAX_AXIS);
  }
  if (!(resolution > 0.0) || !std::isfinite(resolution))
  {
    return FailInvalid("resolution must be positive and finite");
  }
  return VGT_B200_OK;
}

// The min/max keys live in a tiny per-call device allocation so concurrent calls never share
// state; stream-ordered allocation keeps the device entry points asynchronous.
template <typename Key>
struct KeyScratch
{
  Key* ptr = nullptr;
  cudaStream_t stream = nullptr;
  cudaError_t Allocate(cudaStream_t s)
  {
    stream = s;
    return cudaMallocAsync(reinterpret_cast<void**>(&ptr), 2 * sizeof(Key), s);
  }
  ~KeyScratch()
  {
    if (ptr != nullptr)
    {
      cudaFreeAsync(ptr, stream);
    }
  }
};

template <typename In>
int SdfFloatOnDevice(const In* d_in, int64_t nx, int64_t ny, int64_t nz, double resolution,
                     int unknown_is_filled, int add_virtual_border, float* d_sdf,
                     float* d_min_max, cudaStream_t stream)
{
  uint32_t* d_packed = reinterpret_cast<uint32_t*>(d_sdf);  // the output doubles as scratch
  int status = RunLocalPasses<In>(d_in, nx, ny, nz, unknown_is_filled, d_packed, stream);
  if (status != VGT_B200_OK)
  {
    return status;
  }
  KeyScratch<uint32_t> keys;
  if (d_min_max != nullptr)
  {
    VGT_CUDA_TRY(keys.Allocate(stream), "min/max scratch");
  }
  return RunFinalPass<kEmitFloat>(d_packed, nx, ny, nz, 0, ny, resolution, add_virtual_border,
                                  d_sdf, d_min_max, keys.ptr, stream);
}
}  // namespace
}  // namespace edt
}  // namespace vgt_b200

// ================================================================================================
// C-ABI (include/vgt_b200.h)
// ================================================================================================
using namespace vgt_b200;
using namespace vgt_b200::edt;

extern "C"
{
int vgt_b200_sdf_f32_dev(
    const float* d_occupancy, int64_t nx, int64_t ny, int64_t nz, double resolution,
    int unknown_is_filled, int add_virtual_border, int device, float* d_sdf_out, float* d_min_max,
    void* stream)
{
  const int check = CheckSdfArguments(d_occupancy, d_sdf_out, nx, ny, nz, resolution);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  return SdfFloatOnDevice<float>(d_occupancy, nx, ny, nz, resolution, unknown_is_filled,
                                 add_virtual_border, d_sdf_out, d_min_max,
                                 static_cast<cudaStream_t>(stream));
}

int vgt_b200_sdf_f32_dev_profile(
    const float* d_occupancy, int64_t nx, int64_t ny, int64_t nz, double resolution,
    int unknown_is_filled, int add_virtual_border, int device, float* d_sdf_out, float* d_min_max,
    void* stream, float* out_pass_ms)
{
  const int check = CheckSdfArguments(d_occupancy, d_sdf_out, nx, ny, nz, resolution);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  if (out_pass_ms == nullptr)
  {
    return FailInvalid("null out_pass_ms");
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaEvent_t marks[4];
  for (auto& mark : marks)
  {
    VGT_CUDA_TRY(cudaEventCreate(&mark), "cudaEventCreate");
  }
  uint32_t* d_packed = reinterpret_cast<uint32_t*>(d_sdf_out);
  KeyScratch<uint32_t> keys;
  if (d_min_max != nullptr)
  {
    VGT_CUDA_TRY(keys.Allocate(s), "min/max scratch");
  }
  cudaEventRecord(marks[0], s);
  int status = LaunchScan<float>(d_occupancy, d_packed, nx * ny, static_cast<int32_t>(nz),
                                 unknown_is_filled, s);
  cudaEventRecord(marks[1], s);
  if (status == VGT_B200_OK && ny > 1)
  {
    const LineFamily along_y{nx, ny * nz, nz, nz, static_cast<int32_t>(ny)};
    status = LaunchEnvelope<kEmitPacked>(d_packed, d_packed, along_y, Square(nz - 1),
                                         FinalizeParams{}, nullptr, s);
  }
  cudaEventRecord(marks[2], s);
  if (status == VGT_B200_OK)
  {
    status = RunFinalPass<kEmitFloat>(d_packed, nx, ny, nz, 0, ny, resolution,
                                      add_virtual_border, d_sdf_out, d_min_max, keys.ptr, s);
  }
  cudaEventRecord(marks[3], s);
  const cudaError_t sync = cudaStreamSynchronize(s);
  if (status == VGT_B200_OK && sync == cudaSuccess)
  {
    for (int i = 0; i < 3; i++)
    {
      cudaEventElapsedTime(out_pass_ms + i, marks[i], marks[i + 1]);
    }
  }
  for (auto& mark : marks)
  {
    cudaEventDestroy(mark);
  }
  if (status != VGT_B200_OK)
  {
    return status;
  }
  VGT_CUDA_TRY(sync, "profiled SDF generation");
  return VGT_B200_OK;
}

int vgt_b200_sdf_from_mask_f32_dev(
    const uint8_t* d_filled_mask, int64_t nx, int64_t ny, int64_t nz, double resolution,
    int add_virtual_border, int device, float* d_sdf_out, float* d_min_max, void* stream)
{
  const int check = CheckSdfArguments(d_filled_mask, d_sdf_out, nx, ny, nz, resolution);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  return SdfFloatOnDevice<uint8_t>(d_filled_mask, nx, ny, nz, resolution, 0, add_virtual_border,
                                   d_sdf_out, d_min_max, static_cast<cudaStream_t>(stream));
}

int vgt_b200_sdf_f64_dev(
    const float* d_occupancy, int64_t nx, int64_t ny, int64_t nz, double resolution,
    int unknown_is_filled, int add_virtual_border, int device, int32_t* d_scratch,
    double* d_sdf_out, double* d_min_max, void* stream)
{
  const int check = CheckSdfArguments(d_occupancy, d_sdf_out, nx, ny, nz, resolution);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  if (d_scratch == nullptr)
  {
    return FailInvalid("null scratch pointer");
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint32_t* d_packed = reinterpret_cast<uint32_t*>(d_scratch);
  const int status =
      RunLocalPasses<float>(d_occupancy, nx, ny, nz, unknown_is_filled, d_packed, s);
  if (status != VGT_B200_OK)
  {
    return status;
  }
  KeyScratch<unsigned long long> keys;
  if (d_min_max != nullptr)
  {
    VGT_CUDA_TRY(keys.Allocate(s), "min/max scratch");
  }
  return RunFinalPass<kEmitDouble>(d_packed, nx, ny, nz, 0, ny, resolution, add_virtual_border,
                                   d_sdf_out, d_min_max, keys.ptr, s);
}

int vgt_b200_edt_local_passes_dev(
    const float* d_occupancy, int64_t nx_local, int64_t ny, int64_t nz, int unknown_is_filled,
    int device, int32_t* d_out, void* stream)
{
  const int check = CheckSdfArguments(d_occupancy, d_out, nx_local, ny, nz, 1.0);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  return RunLocalPasses<float>(d_occupancy, nx_local, ny, nz, unknown_is_filled,
                               reinterpret_cast<uint32_t*>(d_out),
                               static_cast<cudaStream_t>(stream));
}

int vgt_b200_edt_final_pass_f32_dev(
    const int32_t* d_in, int64_t nx, int64_t ny_local, int64_t nz, int64_t y_offset,
    int64_t ny_total, double resolution, int add_virtual_border, int device, float* d_sdf_out,
    float* d_min_max, void* stream)
{
  const int check = CheckSdfArguments(d_in, d_sdf_out, nx, ny_local, nz, resolution);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  if (y_offset < 0 || ny_total < 1 || ny_total > VGT_B200_MAX_AXIS || y_offset + ny_local > ny_total)
  {
    return FailInvalid("y slab [%lld, %lld) does not fit ny_total = %lld",
                       static_cast<long long>(y_offset),
                       static_cast<long long>(y_offset + ny_local),
                       static_cast<long long>(ny_total));
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  KeyScratch<uint32_t> keys;
  if (d_min_max != nullptr)
  {
    VGT_CUDA_TRY(keys.Allocate(s), "min/max scratch");
  }
  return RunFinalPass<kEmitFloat>(reinterpret_cast<const uint32_t*>(d_in), nx, ny_local, nz,
                                  y_offset, ny_total, resolution, add_virtual_border, d_sdf_out,
                                  d_min_max, keys.ptr, s);
}

// ------------------------------------------------------------------------------------------------
// Host-pointer entry points: allocate, copy in, run, copy out, synchronise.
// ------------------------------------------------------------------------------------------------
}  // extern "C"

namespace
{
template <typename In, typename Out>
int SdfFromHost(const In* h_in, int64_t nx, int64_t ny, int64_t nz, double resolution,
                int unknown_is_filled, int add_virtual_border, int device, Out* h_out,
                Out* out_min, Out* out_max)
{
  const int check = CheckSdfArguments(h_in, h_out, nx, ny, nz, resolution);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  const int64_t count = nx * ny * nz;
  cudaStream_t stream = nullptr;
  VGT_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
  struct StreamGuard
  {
    cudaStream_t s;
    ~StreamGuard() { cudaStreamDestroy(s); }
  } guard{stream};

  DeviceBuffer<In> d_in;
  DeviceBuffer<Out> d_out;
  DeviceBuffer<Out> d_min_max;
  DeviceBuffer<int32_t> d_scratch;
  VGT_CUDA_TRY(d_in.Allocate(count), "cudaMalloc input grid");
  VGT_CUDA_TRY(d_out.Allocate(count), "cudaMalloc SDF");
  VGT_CUDA_TRY(d_min_max.Allocate(2), "cudaMalloc min/max");
  VGT_CUDA_TRY(cudaMemcpyAsync(d_in.get(), h_in, sizeof(In) * count, cudaMemcpyHostToDevice,
                               stream),
               "copy grid to device");
  int status = VGT_B200_OK;
  if constexpr (sizeof(Out) == 4)
  {
    status = SdfFloatOnDevice<In>(d_in.get(), nx, ny, nz, resolution, unknown_is_filled,
                                  add_virtual_border, reinterpret_cast<float*>(d_out.get()),
                                  reinterpret_cast<float*>(d_min_max.get()), stream);
  }
  else
  {
    VGT_CUDA_TRY(d_scratch.Allocate(count), "cudaMalloc scratch");
    status = vgt_b200_sdf_f64_dev(reinterpret_cast<const float*>(d_in.get()), nx, ny, nz,
                                  resolution, unknown_is_filled, add_virtual_border, device,
                                  d_scratch.get(), reinterpret_cast<double*>(d_out.get()),
                                  reinterpret_cast<double*>(d_min_max.get()), stream);
  }
  if (status != VGT_B200_OK)
  {
    cudaStreamSynchronize(stream);
    return status;
  }
  Out min_max[2];
  VGT_CUDA_TRY(cudaMemcpyAsync(h_out, d_out.get(), sizeof(Out) * count, cudaMemcpyDeviceToHost,
                               stream),
               "copy SDF to host");
  VGT_CUDA_TRY(cudaMemcpyAsync(min_max, d_min_max.get(), sizeof(Out) * 2, cudaMemcpyDeviceToHost,
                               stream),
               "copy min/max to host");
  VGT_CUDA_TRY(cudaStreamSynchronize(stream), "SDF generation");
  if (out_min != nullptr)
  {
    *out_min = min_max[0];
  }
  if (out_max != nullptr)
  {
    *out_max = min_max[1];
  }
  return VGT_B200_OK;
}
}  // namespace

extern "C"
{
int vgt_b200_sdf_f32(
    const float* occupancy, int64_t nx, int64_t ny, int64_t nz, double resolution,
    int unknown_is_filled, int add_virtual_border, int device, float* sdf_out, float* out_min,
    float* out_max)
{
  return SdfFromHost<float, float>(occupancy, nx, ny, nz, resolution, unknown_is_filled,
                                   add_virtual_border, device, sdf_out, out_min, out_max);
}

int vgt_b200_sdf_f64(
    const float* occupancy, int64_t nx, int64_t ny, int64_t nz, double resolution,
    int unknown_is_filled, int add_virtual_border, int device, double* sdf_out, double* out_min,
    double* out_max)
{
  return SdfFromHost<float, double>(occupancy, nx, ny, nz, resolution, unknown_is_filled,
                                    add_virtual_border, device, sdf_out, out_min, out_max);
}

int vgt_b200_sdf_from_mask_f32(
    const uint8_t* filled_mask, int64_t nx, int64_t ny, int64_t nz, double resolution,
    int add_virtual_border, int device, float* sdf_out, float* out_min, float* out_max)
{
  return SdfFromHost<uint8_t, float>(filled_mask, nx, ny, nz, resolution, 0, add_virtual_border,
                                     device, sdf_out, out_min, out_max);
}

int vgt_b200_edt_sq_i32(
    const float* occupancy, int64_t nx, int64_t ny, int64_t nz, int unknown_is_filled, int device,
    int32_t* dist_to_filled_sq, int32_t* dist_to_free_sq)
{
  const int check = CheckSdfArguments(occupancy, dist_to_filled_sq, nx, ny, nz, 1.0);
  if (check != VGT_B200_OK)
  {
    return check;
  }
  if (dist_to_free_sq == nullptr)
  {
    return FailInvalid("null output pointer");
  }
  ScopedDevice scoped(device);
  VGT_CUDA_TRY(scoped.Status(), "cudaSetDevice");
  const int64_t count = nx * ny * nz;
  DeviceBuffer<float> d_in;
  DeviceBuffer<uint32_t> d_packed;
  DeviceBuffer<int32_t> d_filled;
  DeviceBuffer<int32_t> d_free;
  VGT_CUDA_TRY(d_in.Allocate(count), "cudaMalloc occupancy");
  VGT_CUDA_TRY(d_packed.Allocate(count), "cudaMalloc packed field");
  VGT_CUDA_TRY(d_filled.Allocate(count), "cudaMalloc field");
  VGT_CUDA_TRY(d_free.Allocate(count), "cudaMalloc field");
  VGT_CUDA_TRY(cudaMemcpy(d_in.get(), occupancy, sizeof(float) * count, cudaMemcpyHostToDevice),
               "copy occupancy to device");
  int status = RunLocalPasses<float>(d_in.get(), nx, ny, nz, unknown_is_filled, d_packed.get(),
                                     nullptr);
  if (status != VGT_B200_OK)
  {
    return status;
  }
  if (nx > 1)
  {
    const LineFamily along_x{1, 0, ny * nz, ny * nz, static_cast<int32_t>(nx)};
    status = LaunchEnvelope<kEmitPacked>(d_packed.get(), d_packed.get(), along_x,
                                         Square(nz - 1) + Square(ny - 1), FinalizeParams{},
                                         nullptr, nullptr);
    if (status != VGT_B200_OK)
    {
      return status;
    }
  }
  const int threads = 256;
  SplitFieldsKernel<<<static_cast<unsigned>((count + threads - 1) / threads), threads>>>(
      d_packed.get(), count, d_filled.get(), d_free.get());
  VGT_CUDA_TRY(cudaGetLastError(), "SplitFieldsKernel launch");
  VGT_CUDA_TRY(cudaMemcpy(dist_to_filled_sq, d_filled.get(), sizeof(int32_t) * count,
                          cudaMemcpyDeviceToHost),
               "copy field to host");
  VGT_CUDA_TRY(cudaMemcpy(dist_to_free_sq, d_free.get(), sizeof(int32_t) * count,
                          cudaMemcpyDeviceToHost),
               "copy field to host");
  return VGT_B200_OK;
}
}  // extern "C"
