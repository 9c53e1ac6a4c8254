// Pass A (the contiguous z axis) for lines of at most 2048 voxels, entirely in registers (sm_100a).
// Included by edt_kernels.cu after edt_device.cuh.
//
// Same algorithm as ScanContiguousAxisVec4Kernel (edt_device.cuh): one warp owns one line, a lane
// owns four consecutive voxels per 128-voxel iteration (128-bit loads and stores), the eight
// lanes that share a 32-voxel word merge their class nibbles into the word, and the nearest
// opposite-class voxel is found with bit scans inside the word and per-word tables outside it.
// The round-1 profile showed that kernel bound by issued instructions (2.4 per voxel, 85 % issue
// slots busy), so this version removes what cost them:
//   * the word is merged with three xor-shuffles instead of a partial-mask redux (which the
//     compiler serialises into one redux per 8-lane group);
//   * the words stay in registers (a lane keeps its group's word of every iteration) and the
//     per-word tables live in the registers of lane w = word index, read back with shuffles:
//     no shared memory, no __syncwarp;
//   * a lane searches twice (left of its first voxel, right of its last), not eight times: inside
//     its four voxels the distances follow by +1 / restart at 1 where the class flips; and the
//     searches are selects over both candidates, not branches that 1-6 lanes take.
// Replaces, for both fields at once: the marking loop (sdfgen.hpp:57-74) and the Z-axis loop of
// ComputeDistanceFieldTransformInPlace (sdfgen.cpp:354-390).
#pragma once

#include "edt_device.cuh"

namespace vgt_b200
{
namespace edt
{
namespace
{
// kIterations = ceil(length / 128) rounded up to 1, 2, 4, 8 or 16. Up to 8 iterations a line has
// at most 32 words and lane w owns word w; with 16 iterations (lines of 1025 .. 2048 voxels) a
// lane owns two words, w and w + 32 ("halves"): each half is scanned on its own and the totals
// cross over (the last filled / free voxel of the first half precedes every word of the second,
// the first one of the second half follows every word of the first).
template <typename Source, int kIterations>
__global__ void __launch_bounds__(kScanWarpsPerBlock* kWarp, (kIterations > 8) ? 4 : 8)
    ScanContiguousAxisRegistersKernel(
    const typename Source::Vector* __restrict__ in, uint4* __restrict__ out, int64_t num_lines,
    int32_t length, int unknown_is_filled)
{
  using Vector = typename Source::Vector;
  constexpr unsigned kFull = 0xffffffffu;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t line = static_cast<int64_t>(blockIdx.x) * kScanWarpsPerBlock + warp;
  if (line >= num_lines)
  {
    return;  // warp-uniform
  }
  const int vectors = length >> 2;  // per line
  const int num_words = (length + 31) >> 5;
  const Vector* src = in + line * vectors;
  uint4* dst = out + line * vectors;
  const int group = lane >> 3;        // which of the 4 words of an iteration
  const int bit0 = (lane & 7) << 2;   // first of this lane's 4 bits in that word

  // 1. classify: 4 voxels per lane; every lane ends up with the whole word of its group.
  uint32_t words[kIterations];
#pragma unroll
  for (int it = 0; it < kIterations; it++)
  {
    const int vector_index = (it << 5) + lane;
    uint32_t nibble = 0;
    if (vector_index < vectors)
    {
      nibble = Source::Nibble(__ldcs(src + vector_index), unknown_is_filled);
    }
    uint32_t word = nibble << bit0;
    word |= __shfl_xor_sync(kFull, word, 1);
    word |= __shfl_xor_sync(kFull, word, 2);
    word |= __shfl_xor_sync(kFull, word, 4);
    words[it] = word;
  }

  // 2. lane w owns word w (= iteration w / 4, group w % 4) of every half: position of the last
  //    filled / free voxel before the word and of the first one after it, by warp scans over the
  //    words.
  constexpr int kHalves = (kIterations + 7) / 8;
  static_assert(kHalves <= 2, "at most 64 words per line");
  uint32_t mine[kHalves];
#pragma unroll
  for (int h = 0; h < kHalves; h++)
  {
    mine[h] = 0;
  }
#pragma unroll
  for (int it = 0; it < kIterations; it++)
  {
    const uint32_t from_leader = __shfl_sync(kFull, words[it], (lane & 3) << 3);
    mine[it >> 3] = ((lane >> 2) == (it & 7)) ? from_leader : mine[it >> 3];
  }
  int before_filled[kHalves], before_free[kHalves], after_filled[kHalves], after_free[kHalves];
  int last_filled[kHalves], last_free[kHalves], first_filled[kHalves], first_free[kHalves];
#pragma unroll
  for (int h = 0; h < kHalves; h++)
  {
    const int w = (h << 5) + lane;
    const uint32_t valid = (w < num_words) ? ValidBits(w, length) : 0u;
    const uint32_t filled_bits = mine[h] & valid;
    const uint32_t free_bits = ~mine[h] & valid;
    last_filled[h] = filled_bits ? (w << 5) + 31 - __clz(filled_bits) : -kFar;
    last_free[h] = free_bits ? (w << 5) + 31 - __clz(free_bits) : -kFar;
    first_filled[h] = filled_bits ? (w << 5) + __ffs(filled_bits) - 1 : kFar;
    first_free[h] = free_bits ? (w << 5) + __ffs(free_bits) - 1 : kFar;
#pragma unroll
    for (int offset = 1; offset < kWarp; offset <<= 1)
    {
      const int up_filled = __shfl_up_sync(kFull, last_filled[h], offset);
      const int up_free = __shfl_up_sync(kFull, last_free[h], offset);
      const int down_filled = __shfl_down_sync(kFull, first_filled[h], offset);
      const int down_free = __shfl_down_sync(kFull, first_free[h], offset);
      if (lane >= offset)
      {
        last_filled[h] = max(last_filled[h], up_filled);
        last_free[h] = max(last_free[h], up_free);
      }
      if (lane + offset < kWarp)
      {
        first_filled[h] = min(first_filled[h], down_filled);
        first_free[h] = min(first_free[h], down_free);
      }
    }
  }
  // what precedes the first word of a half / follows its last word
  int preceding_filled = -kFar, preceding_free = -kFar;
#pragma unroll
  for (int h = 0; h < kHalves; h++)
  {
    int following_filled = kFar, following_free = kFar;
    if (h + 1 < kHalves)
    {
      // (the suffix minimum at lane 0 of the next half covers that whole half)
      following_filled = __shfl_sync(kFull, first_filled[h + 1], 0);
      following_free = __shfl_sync(kFull, first_free[h + 1], 0);
    }
    before_filled[h] = __shfl_up_sync(kFull, max(last_filled[h], preceding_filled), 1);
    before_free[h] = __shfl_up_sync(kFull, max(last_free[h], preceding_free), 1);
    after_filled[h] = __shfl_down_sync(kFull, min(first_filled[h], following_filled), 1);
    after_free[h] = __shfl_down_sync(kFull, min(first_free[h], following_free), 1);
    if (lane == 0)
    {
      before_filled[h] = preceding_filled;
      before_free[h] = preceding_free;
    }
    if (lane == kWarp - 1)
    {
      after_filled[h] = following_filled;
      after_free[h] = following_free;
    }
    // (the prefix maximum at lane 31 covers this whole half)
    preceding_filled = max(preceding_filled, __shfl_sync(kFull, last_filled[h], kWarp - 1));
    preceding_free = max(preceding_free, __shfl_sync(kFull, last_free[h], kWarp - 1));
  }

  // 3. per voxel: nearest opposite-class voxel inside the word (bit scan) or outside (tables).
#pragma unroll
  for (int it = 0; it < kIterations; it++)
  {
    const int vector_index = (it << 5) + lane;
    const int w = (it << 2) + group;
    const int word_start = w << 5;
    // positions relative to the start of the word (all lanes shuffle; stores are predicated)
    // (the half, it >> 3, is known at compile time: the loop is unrolled)
    const int left_of_filled = __shfl_sync(kFull, before_free[it >> 3], w & 31) - word_start;
    const int left_of_free = __shfl_sync(kFull, before_filled[it >> 3], w & 31) - word_start;
    const int right_of_filled = __shfl_sync(kFull, after_free[it >> 3], w & 31) - word_start;
    const int right_of_free = __shfl_sync(kFull, after_filled[it >> 3], w & 31) - word_start;
    const uint32_t word = words[it];
    const uint32_t valid_here = ValidBits(w, length);
    const uint32_t opposite_of_filled = ~word & valid_here;
    const uint32_t opposite_of_free = word & valid_here;
    // Two searches per lane instead of eight: the nearest opposite-class voxel to the LEFT of
    // the lane's first voxel and to the RIGHT of its last one. Inside the lane's four voxels the
    // distances follow by +1 while the class stays the same and restart at 1 where it flips.
    const uint32_t nibble = (word >> bit0) & 0xfu;
    const uint32_t flips = nibble ^ (nibble >> 1);  // bit k: voxels k and k + 1 differ
    const bool first_filled = (nibble & 1u) != 0;
    const bool last_filled = (nibble & 8u) != 0;
    const uint32_t below = (first_filled ? opposite_of_filled : opposite_of_free)
        & ((1u << bit0) - 1u);
    const uint32_t above =
        ((last_filled ? opposite_of_filled : opposite_of_free) >> (bit0 + 3)) >> 1;
    const int left_outside = first_filled ? left_of_filled : left_of_free;
    const int right_outside = last_filled ? right_of_filled : right_of_free;
    int left[4];
    int right[4];
    left[0] = bit0 - ((below != 0) ? (31 - __clz(below)) : left_outside);
    right[3] = ((above != 0) ? (bit0 + 3 + __ffs(above)) : right_outside) - (bit0 + 3);
#pragma unroll
    for (int k = 1; k < 4; k++)
    {
      left[k] = ((flips >> (k - 1)) & 1u) ? 1 : left[k - 1] + 1;
    }
#pragma unroll
    for (int k = 2; k >= 0; k--)
    {
      right[k] = ((flips >> k) & 1u) ? 1 : right[k + 1] + 1;
    }
    uint32_t results[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
      const int nearest = min(left[k], right[k]);
      const uint32_t squared =
          (nearest >= kFarThreshold) ? kNone : static_cast<uint32_t>(nearest * nearest);
      results[k] = (((nibble >> k) & 1u) << 31) | squared;
    }
    if (vector_index < vectors)
    {
      dst[vector_index] = make_uint4(results[0], results[1], results[2], results[3]);
    }
  }
}
}  // namespace
}  // namespace edt
}  // namespace vgt_b200
