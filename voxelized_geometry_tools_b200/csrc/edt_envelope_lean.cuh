// The strided-axis envelope kernel (sm_100a): axes up to 8192 voxels.
// Included by edt_kernels.cu after edt_device.cuh.
//
// One lane owns one line, a warp owns 32 z-adjacent lines (every row access of the warp is one
// 128-byte segment), the Felzenszwalb-Huttenlocher stack of a line lives IN PLACE in the rows of
// the line that were already consumed. Axes of at most 1024 voxels whose partial distances stay
// below 2^21 pack an entry into the row itself (f << 10 | v); longer axes / larger distances keep
// f in the row and v in a uint16 side array addressed exactly like the grid (kSplit, 2 more
// bytes per voxel of scratch). The first version of this kernel was bound by issued
// instructions (~280 warp instructions per row of 32 voxels), a third of them in branches taken
// by 1-4 lanes (run boundaries, sweep advances) and another third in fixed per-row bookkeeping.
// Here:
//   * every address is one IMAD.WIDE.U32 (32-bit row index x 32-bit byte stride + 64-bit base);
//   * phase 1 has ONE pop loop: at a class change the incoming site is the zero-height site of
//     the boundary, otherwise the voxel's own site; the change-only work is a short branch;
//   * phase 1 records, per 32-row word, where the run covering the end of that word ends, so
//     phase 2 finds the end of a run with one bit scan or one table read instead of a search;
//   * phase 2 takes the zero site right of the run as a min() in the output expression, so the
//     sweep only ever walks stored sites and the advance body is a load and a decode;
//   * phase 2 walks word by word (class word in a register, no per-row shared-memory logic);
//   * when (largest h) x (length) < 2^31 the pop test is done in 32-bit arithmetic;
//   * send-layout output and the virtual border are template flags, out of the common path.
//
// Replaces the X / Y loops of ComputeDistanceFieldTransformInPlace (sdfgen.cpp:276-351) and the
// 1-D transforms (sdfgen.cpp:85-226) for both fields at once; in finalize mode also the combine
// loop (sdfgen.hpp:85-108) and Lock()'s min/max (sdf.hpp:765-787).
//
// The input buffer is destroyed; the output must be a different buffer.
#pragma once

#include "edt_device.cuh"

namespace vgt_b200
{
namespace edt
{
namespace
{
// One warp per tile of 32 adjacent lines, four tiles per block.
constexpr int kLineWarpsPerBlock = 4;
// Packed stack entries (f << 10 | v): axes up to 1024 voxels, partial distances below 2^21.
constexpr int kInPlacePositionBits = 10;
constexpr int kInPlaceMaxLength = 1 << kInPlacePositionBits;
constexpr int64_t kInPlaceMaxInput = (int64_t{1} << (31 - kInPlacePositionBits)) - 1;

// crossing(below, middle) >= crossing(middle, above), cross-multiplied; see MiddleIsHidden.
template <bool kNarrow>
__device__ __forceinline__ bool HiddenTest(int32_t below_v, int32_t below_h, int32_t middle_v,
                                           int32_t middle_h, int32_t above_v, int32_t above_h)
{
  if constexpr (kNarrow)
  {
    // |h differences| * (position differences) < 2^31 is guaranteed by the launcher.
    return (middle_h - below_h) * (above_v - middle_v) >= (above_h - middle_h) * (middle_v - below_v);
  }
  else
  {
    return static_cast<long long>(middle_h - below_h) * static_cast<long long>(above_v - middle_v)
        >= static_cast<long long>(above_h - middle_h) * static_cast<long long>(middle_v - below_v);
  }
}

// Per-launch scratch in global memory (no shared memory at all, so occupancy is set by registers
// only and the whole L1 serves the stack accesses): class words [num_words][lines] (uint32, bit b
// of word w of a line = class of row 32 w + b) followed by the run-end table [num_words][lines]
// (uint16): entry w of a line = first row after the end of word w whose class differs from the
// class of the word's last row (`length` when there is none). A warp reads / writes 32 adjacent
// lines of one word: one 128-byte (64-byte) segment per 32 rows.
__host__ __device__ inline size_t LeanScratchBytes(int length, int64_t lines)
{
  const int64_t num_words = (length + 31) >> 5;
  return static_cast<size_t>(num_words * lines) * (sizeof(uint32_t) + sizeof(uint16_t));
}

// A zero site that is "not there": far enough that its squared distance is >= kNone for every
// q < 8192, near enough that the square still fits 32 bits.
constexpr int32_t kFarZeroSite = 46341 + VGT_B200_MAX_AXIS;
// Height of a stored site that must not be taken (it belongs to a later run): above
// kNoSiteHeight even after the -2*v*q term, so it never beats an absent winner.
constexpr int32_t kBlockedHeight = 0x5fffffff;
// Position of "no stored site left": above every row index, small enough that the -2*v*q term
// of the comparison cannot overflow or pull kBlockedHeight below kNoSiteHeight.
constexpr int32_t kAbsentPosition = 0x4000;

// A stored site as read back from the stack: position and f (h = f + v * v).
struct StackEntry
{
  int32_t v;
  int32_t f;
};

// kBlocksPerSm: resident blocks (of 4 warps) the register budget is cut for. The finalizing modes
// run at 8 (64 registers, no spills: faster than 10 or 12 blocks with spills in the hot loops).
// The packed mode is built for 12 (40 registers, no spills) and for 16 (32 registers, a few
// spilled words); the launcher takes 16 only when it saves a whole wave (e.g. 8192 warp-tiles on
// 148 SMs: one wave at 64 warps per SM, 1.15 at 48).
template <int kMode, bool kNarrow, bool kSend, bool kBorder, bool kSplit, int kBlocksPerSm>
__global__ void __launch_bounds__(kLineWarpsPerBlock* kWarp, kBlocksPerSm)
    EnvelopeAxisLeanKernel(uint32_t* in, typename OutputOf<kMode>::Type* out, uint16_t* positions,
                           uint32_t* class_scratch, LineFamily family, FinalizeParams finalize,
                           typename OutputOf<kMode>::Key* min_max_keys,
                           const uint32_t* redo_list)
{
  using Out = typename OutputOf<kMode>::Type;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int length = family.length;
  const int num_words = static_cast<int>(family.num_words);

  // (the launcher guarantees that tile, line and column counts fit 31 bits)
  const uint32_t tiles_per_outer = static_cast<uint32_t>((family.inner_count + kWarp - 1) / kWarp);
  uint32_t tile_index = blockIdx.x * kLineWarpsPerBlock + warp;
  if (redo_list != nullptr)
  {
    // Launch after the window kernel (edt_envelope_window.cuh; buffer layout there): either
    // only the tiles it gave up on (word 0 = how many, words 4.. = their indices), or, when its
    // pilot found the map deep (word 1), every tile.
    if (redo_list[1] == 0u)
    {
      if (tile_index >= redo_list[0])
      {
        return;  // warp-uniform
      }
      tile_index = redo_list[4 + tile_index];
    }
  }
  if (tile_index >= tiles_per_outer * static_cast<uint32_t>(family.num_outer))
  {
    return;  // warp-uniform
  }
  const uint32_t outer = tile_index / tiles_per_outer;
  const uint32_t column = (tile_index - outer * tiles_per_outer) * kWarp + lane;
  const bool active = column < static_cast<uint32_t>(family.inner_count);
  const int64_t first = static_cast<int64_t>(outer) * family.outer_stride + column;
  const uint32_t stride_bytes = family.stride_bytes;
  // scratch of this line: class word w at class_words + w * lines, table entry w likewise
  const int64_t lines = family.num_outer * family.inner_count;
  const uint32_t word_stride_bytes = static_cast<uint32_t>(lines) * 4u;
  const uint32_t line_index = outer * static_cast<uint32_t>(family.inner_count) + column;
  char* const class_words = reinterpret_cast<char*>(class_scratch + line_index);
  const auto class_word_address = [&](uint32_t w)
  {
    return reinterpret_cast<uint32_t*>(class_words + static_cast<uint64_t>(w) * word_stride_bytes);
  };
  const auto run_end_address = [&](uint32_t w)
  {
    char* const table = reinterpret_cast<char*>(class_scratch + static_cast<int64_t>(num_words) * lines)
        + 2 * static_cast<uint64_t>(line_index);
    return reinterpret_cast<uint16_t*>(table + static_cast<uint64_t>(w) * (word_stride_bytes >> 1));
  };

  Out lane_min = PositiveInfinity<Out>();
  Out lane_max = -PositiveInfinity<Out>();

  if (active)
  {
    // Pinned in a register pair: the compiler otherwise re-derives it from the kernel parameters
    // (ten instructions) at every prefetch inside the hot loop, which is issue-bound.
    char* line = reinterpret_cast<char*>(in + first);
    asm volatile("" : "+l"(line));
    const auto row_address = [&](uint32_t row)
    {
      return reinterpret_cast<uint32_t*>(line + static_cast<uint64_t>(row) * stride_bytes);
    };
    // stack entries: (f << 10 | v) in the row, or f in the row and v in the side array
    char* const line_positions = kSplit ? reinterpret_cast<char*>(positions + first) : nullptr;
    const auto position_address = [&](uint32_t row)
    {
      return reinterpret_cast<uint16_t*>(line_positions
                                         + static_cast<uint64_t>(row) * (stride_bytes >> 1));
    };
    const auto store_entry = [&](uint32_t at_slot, int32_t v, uint32_t f)
    {
      if constexpr (kSplit)
      {
        *row_address(at_slot) = f;
        *position_address(at_slot) = static_cast<uint16_t>(v);
      }
      else
      {
        *row_address(at_slot) = (f << kInPlacePositionBits) | static_cast<uint32_t>(v);
      }
    };
    const auto load_entry = [&](uint32_t at_slot)
    {
      const uint32_t e = *row_address(at_slot);
      if constexpr (kSplit)
      {
        return StackEntry{static_cast<int32_t>(*position_address(at_slot)),
                          static_cast<int32_t>(e)};
      }
      else
      {
        return StackEntry{static_cast<int32_t>(e & (kInPlaceMaxLength - 1)),
                          static_cast<int32_t>(e >> kInPlacePositionBits)};
      }
    };

    // ------------------------------------------------------------------ phase 1: build stacks
    // Entries of the current run: [implicit zero site left of the run (if any)] + stored sites.
    uint32_t slot = 0;    // stored sites of the whole line so far == next free row
    int entries = 0;      // entries of the current run, the implicit left site included
    int32_t left_v = -1;  // position of the zero site left of the run; -1: the run starts the line
    int32_t top_v = 0, top_h = 0, below_v = 0, below_h = 0;
    uint32_t word_accumulator = 0;

    // next_word: the word of row q + 1, or kNone with the class of `word` when there is none
    // (or it is not known yet), which disables the pre-filter below for this row.
    const auto process_row = [&](const int q, const uint32_t word, uint32_t& previous_word,
                                 const uint32_t next_word)
    {
      const uint32_t value = word & kNone;
      const bool change = static_cast<int32_t>(word ^ previous_word) < 0;
      // Pre-filter: a site that lies on or above the segment between its two neighbour sites
      // g(q-1), g(q+1) (g = f + v^2; a neighbour of the other class is a zero site, f = 0) is not
      // a vertex of the lower hull, so pushing it would only make the next row pop it again.
      // About 40 % of the sites of a distance-like field go this way. kNone neighbours make
      // the signed comparison fail, so they never filter.
      const int32_t before = change ? 0 : static_cast<int32_t>(previous_word & kNone);
      const int32_t after = (static_cast<int32_t>(word ^ next_word) < 0)
          ? 0 : static_cast<int32_t>(next_word & kNone);
      const bool on_hull_locally = 2 * static_cast<int32_t>(value) - 2 - before < after;
      previous_word = word;
      // bit b of the accumulator after 32 rows = class of row 32w + b
      word_accumulator = (word_accumulator >> 1) | (word & kClassBit);
      const bool finite = (value != kNone) && (on_hull_locally || q == 0);
      const int32_t qq = q * q;
      const int32_t own_h = static_cast<int32_t>(value) + qq;
      // The site that arrives at this row: the zero-height site that closes the run at a class
      // change (it hides what it hides but is never stored: phase 2 re-creates it from the class
      // bits), else the voxel's own site.
      const int32_t incoming_h = change ? qq : own_h;
      if (change || finite)
      {
        while (entries >= 2 && HiddenTest<kNarrow>(below_v, below_h, top_v, top_h, q, incoming_h))
        {
          entries--;
          slot--;
          top_v = below_v;
          top_h = below_h;
          // The entry under the new top: the implicit left site when the run holds one stored
          // site, else stored entry slot - 2. The load is unconditional (clamped to row 0 when
          // there is nothing to read): a branch around it costs more than the spare load.
          const StackEntry e =
              load_entry(static_cast<uint32_t>(max(static_cast<int32_t>(slot) - 2, 0)));
          const bool from_left = (entries == 2) && (left_v >= 0);
          const int32_t site_v = from_left ? left_v : e.v;
          const int32_t site_f = from_left ? 0 : e.f;
          below_v = site_v;
          below_h = site_f + site_v * site_v;
        }
      }
      if (change)
      {
        // words whose last row the closed run covers: the run that covers them ends at q
#pragma unroll 1
        for (int w = (left_v + 1) >> 5; w < (q >> 5); w++)
        {
          *run_end_address(static_cast<uint32_t>(w)) = static_cast<uint16_t>(q);
        }
        left_v = q - 1;
        top_v = left_v;
        top_h = left_v * left_v;
        entries = 1;
      }
      if (finite)
      {
        // (after a change the run holds one entry, so no pop test is due for the own site)
        store_entry(slot, q, value);
        slot++;
        below_v = top_v;
        below_h = top_h;
        top_v = q;
        top_h = own_h;
        entries++;
      }
    };

    {
      constexpr int kBatch = 2;
      uint32_t next_rows[kBatch];
      const int full = length & ~(kBatch - 1);
      if (full > 0)
      {
#pragma unroll
        for (int u = 0; u < kBatch; u++)
        {
          next_rows[u] = __ldcs(row_address(static_cast<uint32_t>(u)));
        }
      }
      // the first row never is a class change
      uint32_t previous_word = (full > 0) ? next_rows[0] : __ldcs(row_address(0));
      int q0 = 0;
      for (; q0 < full; q0 += kBatch)
      {
        uint32_t rows[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; u++)
        {
          rows[u] = next_rows[u];
        }
        if (q0 + kBatch < full)
        {
          // prefetch the next batch: its rows are above every slot this batch can write
#pragma unroll
          for (int u = 0; u < kBatch; u++)
          {
            next_rows[u] = __ldcs(row_address(static_cast<uint32_t>(q0 + kBatch + u)));
          }
        }
#pragma unroll
        for (int u = 0; u < kBatch; u++)
        {
          // after the last full batch next_rows still holds this batch: no filtering there
          const uint32_t next_word = (u + 1 < kBatch)
              ? rows[u + 1]
              : ((q0 + kBatch < full) ? next_rows[0] : (rows[u] | kNone));
          process_row(q0 + u, rows[u], previous_word, next_word);
        }
        if ((q0 & 31) == 32 - kBatch)
        {
          __stcs(class_word_address(static_cast<uint32_t>(q0 >> 5)), word_accumulator);
        }
      }
#pragma unroll 1
      for (int q = q0; q < length; q++)
      {
        const uint32_t word = __ldcs(row_address(static_cast<uint32_t>(q)));
        process_row(q, word, previous_word, word | kNone);
        if ((q & 31) == 31)
        {
          __stcs(class_word_address(static_cast<uint32_t>(q >> 5)), word_accumulator);
        }
      }
      if ((length & 31) != 0)
      {
        // align the partial last word; the arithmetic shift repeats the class of the last row
        // in the bits past the end of the line, so they never read as a class change
        __stcs(class_word_address(static_cast<uint32_t>(num_words - 1)),
               static_cast<uint32_t>(static_cast<int32_t>(word_accumulator)
                                     >> (32 - (length & 31))));
      }
#pragma unroll 1
      for (int w = (left_v + 1) >> 5; w < num_words; w++)
      {
        *run_end_address(static_cast<uint32_t>(w)) = static_cast<uint16_t>(length);
      }
    }

    // ------------------------------------------------------------------ phase 2: sweep
    // Per run: the winner starts as the zero site left of the run, the sweep walks the run's
    // stored sites, and the zero site right of the run enters as a min() at the output.
    const uint32_t stored_total = slot;
    uint32_t cursor = 0;
    int32_t pending_v = kAbsentPosition;  // the next stored site, decoded
    int32_t pending_h = kNoSiteHeight;
    // The entry after the pending one is already in flight, so an advance never waits for
    // memory unless two advances follow each other closely. Entries written in phase 1 have
    // mostly left L2 by now (all tiles of the grid are in flight at once).
    const uint32_t last_row = family.last_row;
    // (kept undecoded: one register when entries are packed)
    uint32_t following_word = *row_address(min(1u, last_row));
    uint32_t following_position = kSplit ? *position_address(min(1u, last_row)) : 0u;
    const auto decode_entry = [&](const uint32_t word, const uint32_t position)
    {
      if constexpr (kSplit)
      {
        return StackEntry{static_cast<int32_t>(position), static_cast<int32_t>(word)};
      }
      else
      {
        return StackEntry{static_cast<int32_t>(word & (kInPlaceMaxLength - 1)),
                          static_cast<int32_t>(word >> kInPlacePositionBits)};
      }
    };
    const auto decode_pending = [&](const StackEntry e, const bool present)
    {
      pending_v = present ? e.v : kAbsentPosition;
      pending_h = e.f + e.v * e.v;  // unused when absent
    };
    decode_pending(load_entry(0), 0u < stored_total);
    const auto load_pending = [&]()
    {
      // cursor was just incremented: the pending entry is the one that was in flight
      decode_pending(decode_entry(following_word, following_position), cursor < stored_total);
      // (unconditional, clamped to the line: a branch around the load costs more than the load)
      const uint32_t row = min(cursor + 1, last_row);
      following_word = *row_address(row);
      if constexpr (kSplit)
      {
        following_position = *position_address(row);
      }
    };
    int32_t winner_v = 0, winner_h = kNoSiteHeight;
    int32_t candidate_h = kBlockedHeight;  // pending_h if the pending site is in this run
    int32_t right_v = kFarZeroSite;
    int run_end = 0;

    int32_t border_yz = 0x7fffffff;
    if constexpr (kBorder)
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

    // Output rows: row q is at write_origin + q * out_stride_bytes (one IMAD.WIDE per row). In
    // send layout the origin jumps at part boundaries, which are the same for every lane, so
    // that check is warp-uniform.
    char* write_origin = reinterpret_cast<char*>(out + first);
    const uint32_t out_stride_bytes = family.out_stride_bytes;
    int next_part_start = 0;
    uint32_t previous_class = 0;
    int q = 0;
    // class word and table entry of the next word are loaded one word ahead
    uint32_t next_class_word = __ldcs(class_word_address(0));
    uint32_t next_run_end = __ldcs(run_end_address(0));
    uint32_t class_word = 0;
    int run_end_after_word = 0;
    const auto next_word = [&](const int w)
    {
      class_word = next_class_word;
      run_end_after_word = static_cast<int>(next_run_end);
      const uint32_t ahead = static_cast<uint32_t>(min(w + 1, num_words - 1));
      next_class_word = __ldcs(class_word_address(ahead));
      next_run_end = __ldcs(run_end_address(ahead));
    };
    // one row of the sweep; b = q % 32
    const auto sweep_row = [&](const int b)
      {
        if constexpr (kSend)
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
            Out* part_row;
            if (family.scatter_base[0] != nullptr)
            {
              // this part goes straight to its owner's receive buffer over NVLink
              const int part = (q < wide_rows)
                  ? (q / wide)
                  : (family.out_extra + (q - wide_rows) / family.out_base);
              // (selected with compares: indexing the kernel-parameter array with a register
              // would force a local-memory copy of the whole parameter struct)
              uint32_t* base = family.scatter_base[0];
#pragma unroll
              for (int i = 1; i < 8; i++)
              {
                base = (part == i) ? family.scatter_base[i] : base;
              }
              part_row = reinterpret_cast<Out*>(base)
                  + family.inner_count * ((family.scatter_row_offset + outer) * rows + (q - y0))
                  + column;
            }
            else
            {
              part_row = out + family.inner_count * (family.num_outer * y0 + outer * rows + (q - y0))
                  + column;
            }
            write_origin = reinterpret_cast<char*>(part_row)
                - static_cast<uint64_t>(static_cast<uint32_t>(q)) * out_stride_bytes;
          }
        }
        const uint32_t remaining_bits = class_word >> b;  // bit 0 = class of row q
        const uint32_t filled = remaining_bits & 1u;
        if (filled != previous_class || q == 0)
        {
          // A run starts at q. Its end: the next opposite-class bit of this word, else the table
          // (bits past the end of the line repeat the last class, so they never end a run).
          const uint32_t different = filled ? ~remaining_bits & (0xffffffffu >> b) : remaining_bits;
          run_end = different ? q + __ffs(different) - 1 : run_end_after_word;
          // Drop stored sites of earlier runs that the sweep never reached.
          while (pending_v < q)
          {
            cursor++;
            load_pending();
          }
          winner_v = (q > 0) ? q - 1 : 0;
          winner_h = (q > 0) ? (q - 1) * (q - 1) : kNoSiteHeight;
          right_v = (run_end < length) ? run_end : kFarZeroSite;
          candidate_h = (pending_v < run_end) ? pending_h : kBlockedHeight;
          previous_class = filled;
        }

        // Advance while the next stored site of the run is strictly lower at q (F-H
        // "while z[k+1] < q"). An absent winner carries kNoSiteHeight and loses to any real
        // site; a site of a later run carries kBlockedHeight and never wins.
        const int32_t minus_two_q = -2 * q;
        while (candidate_h + pending_v * minus_two_q < winner_h + winner_v * minus_two_q)
        {
          winner_v = pending_v;
          winner_h = candidate_h;
          cursor++;
          load_pending();
          candidate_h = (pending_v < run_end) ? pending_h : kBlockedHeight;
        }

        // squared distance: the stored / left winner, or the zero site right of the run
        uint32_t squared = (winner_h != kNoSiteHeight)
            ? static_cast<uint32_t>(winner_h + winner_v * minus_two_q + q * q)
            : kNone;
        const uint32_t to_right = static_cast<uint32_t>(right_v - q);
        squared = min(min(squared, to_right * to_right), kNone);

        char* const write_at =
            write_origin + static_cast<uint64_t>(static_cast<uint32_t>(q)) * out_stride_bytes;
        if constexpr (kMode == kEmitPacked)
        {
          __stcs(reinterpret_cast<uint32_t*>(write_at), (remaining_bits << 31) | squared);
        }
        else
        {
          if constexpr (kBorder)
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
          const Out value = SignedDistanceFromTable<Out>(
              filled, squared, finalize.resolution,
              static_cast<const Out*>(finalize.magnitude_table), finalize.magnitude_table_size);
          __stcs(reinterpret_cast<Out*>(write_at), value);
          lane_min = (value < lane_min) ? value : lane_min;
          lane_max = (value > lane_max) ? value : lane_max;
        }
        q++;
      };
    // full words with a constant trip count, then the partial last word
    const int full_words = length >> 5;
    for (int w = 0; w < full_words; w++)
    {
      next_word(w);
#pragma unroll 1
      for (int b = 0; b < 32; b++)
      {
        sweep_row(b);
      }
    }
    if ((length & 31) != 0)
    {
      next_word(full_words);
      const int tail_rows = length & 31;
#pragma unroll 1
      for (int b = 0; b < tail_rows; b++)
      {
        sweep_row(b);
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
}  // namespace
}  // namespace edt
}  // namespace vgt_b200
