// The fast strided-axis envelope kernel (sm_100a). Included by edt_kernels.cu after edt_device.cuh.
//
// For axes of at most 1024 voxels whose finite inputs stay below 2^21 the envelope stack is not in
// shared memory but IN PLACE in the line's own global storage: stack slot k of a line lives in row
// k of that line. Slot k <= current position always holds (one entry per voxel at most), so a slot
// only ever overwrites an input row that has already been consumed. Shared memory then holds just
// the class bitmask (length/8 bytes per line), occupancy is set by registers rather than by the
// stack, and the dependent-latency chains of many warps overlap.
//
// The input buffer is destroyed; the output must be a different buffer.
//
// Work split: one lane owns one line; a warp owns 32 adjacent lines so that every global access of
// the warp (input row, stack slot row, output row) is a contiguous 128-byte segment when lanes are
// at the same row. Replaces the X / Y loops of ComputeDistanceFieldTransformInPlace
// (sdfgen.cpp:276-351) and the 1-D transforms (sdfgen.cpp:85-226) for both fields at once; in
// finalize mode also the combine loop (sdfgen.hpp:85-108) and Lock()'s min/max (sdf.hpp:765-787).
#pragma once

#include "edt_device.cuh"

namespace vgt_b200
{
namespace edt
{
namespace
{
constexpr int kLineWarpsPerBlock = 4;
constexpr int kInPlacePositionBits = 10;
constexpr int kInPlaceMaxLength = 1 << kInPlacePositionBits;
constexpr int64_t kInPlaceMaxInput = (int64_t{1} << (31 - kInPlacePositionBits)) - 1;

__device__ __forceinline__ uint32_t PackInPlace(int32_t v, uint32_t f)
{
  return (f << kInPlacePositionBits) | static_cast<uint32_t>(v);
}

__device__ __forceinline__ Site UnpackInPlace(uint32_t e)
{
  const int32_t v = static_cast<int32_t>(e & (kInPlaceMaxLength - 1));
  const int32_t f = static_cast<int32_t>(e >> kInPlacePositionBits);
  return Site{v, f + v * v};
}

// kSplit = false: one packed 32-bit entry (f << 10 | v) per slot, for axes <= 1024 voxels and
//                  partial distances < 2^21.
// kSplit = true:  f stays a full 31-bit word in place and v goes to `positions`, a uint16 side
//                  array addressed exactly like the grid (2 more bytes per voxel of scratch); this
//                  lifts the limits to 8192 voxels per axis and 2^31 - 1.
template <int kMode, bool kSplit>
__global__ void __launch_bounds__(kLineWarpsPerBlock* kWarp, (kMode == kEmitPacked) ? 16 : 12)
    EnvelopeAxisInPlaceStackKernel(
    uint32_t* in, typename OutputOf<kMode>::Type* out, uint16_t* positions, LineFamily family,
    FinalizeParams finalize, typename OutputOf<kMode>::Key* min_max_keys)
{
  using Out = typename OutputOf<kMode>::Type;
  extern __shared__ uint32_t class_smem[];  // [warp][word][lane]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int length = family.length;
  const int num_words = (length + 31) >> 5;
  uint32_t* class_words = class_smem + static_cast<size_t>(warp) * num_words * kWarp;

  const int64_t tiles_per_outer = (family.inner_count + kWarp - 1) / kWarp;
  const int64_t tile_index = static_cast<int64_t>(blockIdx.x) * kLineWarpsPerBlock + warp;
  if (tile_index >= tiles_per_outer * family.num_outer)
  {
    return;  // warp-uniform
  }
  const int64_t outer = tile_index / tiles_per_outer;
  const int64_t tile = tile_index - outer * tiles_per_outer;
  const int64_t column = tile * kWarp + lane;
  const bool active = column < family.inner_count;
  const int64_t first = outer * family.outer_stride + column;
  const int64_t stride = family.line_stride;

  Out lane_min = PositiveInfinity<Out>();
  Out lane_max = -PositiveInfinity<Out>();

  if (active)
  {
    uint32_t* line = in + first;
    uint16_t* line_positions = kSplit ? (positions + first) : nullptr;
    const auto store_entry = [&](int at_slot, int32_t v, uint32_t f)
    {
      const int64_t offset = static_cast<int64_t>(at_slot) * stride;
      if constexpr (kSplit)
      {
        line[offset] = f;
        line_positions[offset] = static_cast<uint16_t>(v);
      }
      else
      {
        line[offset] = PackInPlace(v, f);
      }
    };
    const auto load_entry = [&](int at_slot)
    {
      const int64_t offset = static_cast<int64_t>(at_slot) * stride;
      if constexpr (kSplit)
      {
        const int32_t v = static_cast<int32_t>(line_positions[offset]);
        return Site{v, static_cast<int32_t>(line[offset]) + v * v};
      }
      else
      {
        return UnpackInPlace(line[offset]);
      }
    };

    // ------------------------------------------------------------------ phase 1: build stacks
    int slot = 0;   // next free stack slot == row of the line it will be written to
    int depth = 0;  // stored sites of the current run
    bool has_left = false;
    Site left_zero{0, 0};
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
          below = load_entry(slot - 2);
        }
        else
        {
          below = left_zero;
        }
      }
    };

    constexpr int kBatch = 4;
    for (int q0 = 0; q0 < length; q0 += kBatch)
    {
      uint32_t batch[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; u++)
      {
        batch[u] = (q0 + u < length) ? __ldcg(line + static_cast<int64_t>(q0 + u) * stride) : 0u;
      }
#pragma unroll
      for (int u = 0; u < kBatch; u++)
      {
        const int q = q0 + u;
        if (q < length)
        {
          const uint32_t word = batch[u];
          const uint32_t filled = word >> 31;
          const uint32_t value = word & kNone;
          word_accumulator |= filled << (q & 31);
          if ((q & 31) == 31 || q == length - 1)
          {
            class_words[(q >> 5) * kWarp + lane] = word_accumulator;
            word_accumulator = 0;
          }
          if (q > 0 && filled != previous_class)
          {
            // The run ends: voxel q is a zero-height site for it. It hides what it hides but is
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
            store_entry(slot, q, value);
            below = top;
            top = incoming;
            depth++;
            slot++;
          }
        }
      }
    }

    // ------------------------------------------------------------------ phase 2: sweep
    const int stored_total = slot;
    int cursor = 0;
    const auto load_stored = [&](int index)
    {
      if (index < stored_total)
      {
        return load_entry(index);
      }
      return Site{kNoSitePosition, kNoSiteHeight};
    };
    Site pending = load_stored(0);
    Site winner{0, kNoSiteHeight};
    Site right_zero{0, kNoSiteHeight};
    int run_end = 0;
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

    // Output row pointer. In send layout (family.out_parts > 0) it jumps at part boundaries,
    // which are the same for every lane, so the check below is warp-uniform.
    Out* write_at = out + first;
    int next_part_start = length;
    if (family.out_parts > 0)
    {
      next_part_start = 0;
    }
    for (int q = 0; q < length; q++)
    {
      if (q == next_part_start)
      {
        // Part h covers rows [y0, y0 + rows); its block starts at inner * num_outer * y0.
        const int wide = family.out_base + 1;
        const int wide_rows = family.out_extra * wide;
        int y0;
        int rows;
        if (q < wide_rows)
        {
          y0 = (q / wide) * wide;
          rows = wide;
        }
        else
        {
          y0 = wide_rows + ((q - wide_rows) / family.out_base) * family.out_base;
          rows = family.out_base;
        }
        next_part_start = y0 + rows;
        if (family.scatter_base[0] != nullptr)
        {
          // this part goes straight to its owner's receive buffer over NVLink
          const int part = (q < wide_rows) ? (q / wide)
                                           : (family.out_extra + (q - wide_rows) / family.out_base);
          // (selected with compares: indexing the kernel-parameter array with a register would
          // force a local-memory copy of the whole parameter struct)
          uint32_t* base = family.scatter_base[0];
#pragma unroll
          for (int i = 1; i < 8; i++)
          {
            base = (part == i) ? family.scatter_base[i] : base;
          }
          write_at = reinterpret_cast<Out*>(base)
              + family.inner_count * ((family.scatter_row_offset + outer) * rows + (q - y0))
              + column;
        }
        else
        {
          write_at = out + family.inner_count * (family.num_outer * y0 + outer * rows + (q - y0))
              + column;
        }
      }
      if ((q & 31) == 0)
      {
        class_word = class_words[(q >> 5) * kWarp + lane];
      }
      const uint32_t filled = (class_word >> (q & 31)) & 1u;
      if (q == 0 || filled != previous_class)
      {
        // A run starts at q: find where it ends from the class bits.
        int w = q >> 5;
        uint32_t different = (filled ? ~class_word : class_word) & (0xffffffffu << (q & 31));
        while (different == 0 && ++w < num_words)
        {
          const uint32_t bits = class_words[w * kWarp + lane];
          different = filled ? ~bits : bits;
        }
        run_end = different ? min(length, (w << 5) + __ffs(different) - 1) : length;
        // Drop stored sites of earlier runs that the sweep never reached.
        while (pending.v < q)
        {
          cursor++;
          pending = load_stored(cursor);
        }
        right_zero =
            (run_end < length) ? Site{run_end, run_end * run_end} : Site{0, kNoSiteHeight};
        winner = (q > 0) ? Site{q - 1, (q - 1) * (q - 1)} : Site{0, kNoSiteHeight};
      }
      previous_class = filled;

      // Advance while the next candidate is strictly lower at q (F-H "while z[k+1] < q").
      // An absent candidate / winner carries kNoSiteHeight and so never wins / always loses.
      while (true)
      {
        const bool from_stack = pending.v < run_end;
        const Site candidate = from_stack ? pending : right_zero;
        const int32_t candidate_value = candidate.h - 2 * candidate.v * q;
        const int32_t winner_value = winner.h - 2 * winner.v * q;
        if (!(candidate_value < winner_value))
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
          right_zero = Site{0, kNoSiteHeight};
        }
      }

      uint32_t squared = kNone;
      if (winner.h != kNoSiteHeight)
      {
        squared = static_cast<uint32_t>(winner.h - 2 * winner.v * q + q * q);
      }

      if constexpr (kMode == kEmitPacked)
      {
        *write_at = (filled << 31) | squared;
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
        *write_at = value;
        lane_min = (value < lane_min) ? value : lane_min;
        lane_max = (value > lane_max) ? value : lane_max;
      }
      write_at += stride;
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
}  // namespace
}  // namespace edt
}  // namespace vgt_b200
