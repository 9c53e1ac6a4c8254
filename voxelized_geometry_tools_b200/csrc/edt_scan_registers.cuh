// Pass A (the contiguous z axis) for lines of at most 1024 voxels, entirely in registers (sm_100a).
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
// kIterations = ceil(length / 128) rounded up to 1, 2, 4 or 8 (so at most 32 words per line).
template <typename Source, int kIterations>
__global__ void __launch_bounds__(kScanWarpsPerBlock* kWarp, 8) ScanContiguousAxisRegistersKernel(
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

  // 2. lane w owns word w (= iteration w / 4, group w % 4): position of the last filled / free
  //    voxel before the word and of the first one after it, by warp scans over the words.
  uint32_t mine = 0;
#pragma unroll
  for (int it = 0; it < kIterations; it++)
  {
    const uint32_t from_leader = __shfl_sync(kFull, words[it], (lane & 3) << 3);
    mine = ((lane >> 2) == it) ? from_leader : mine;
  }
  const uint32_t valid = (lane < num_words) ? ValidBits(lane, length) : 0u;
  const uint32_t filled_bits = mine & valid;
  const uint32_t free_bits = ~mine & valid;
  int last_filled = filled_bits ? (lane << 5) + 31 - __clz(filled_bits) : -kFar;
  int last_free = free_bits ? (lane << 5) + 31 - __clz(free_bits) : -kFar;
  int first_filled = filled_bits ? (lane << 5) + __ffs(filled_bits) - 1 : kFar;
  int first_free = free_bits ? (lane << 5) + __ffs(free_bits) - 1 : kFar;
#pragma unroll
  for (int offset = 1; offset < kWarp; offset <<= 1)
  {
    const int up_filled = __shfl_up_sync(kFull, last_filled, offset);
    const int up_free = __shfl_up_sync(kFull, last_free, offset);
    const int down_filled = __shfl_down_sync(kFull, first_filled, offset);
    const int down_free = __shfl_down_sync(kFull, first_free, offset);
    if (lane >= offset)
    {
      last_filled = max(last_filled, up_filled);
      last_free = max(last_free, up_free);
    }
    if (lane + offset < kWarp)
    {
      first_filled = min(first_filled, down_filled);
      first_free = min(first_free, down_free);
    }
  }
  int before_filled = __shfl_up_sync(kFull, last_filled, 1);
  int before_free = __shfl_up_sync(kFull, last_free, 1);
  int after_filled = __shfl_down_sync(kFull, first_filled, 1);
  int after_free = __shfl_down_sync(kFull, first_free, 1);
  if (lane == 0)
  {
    before_filled = -kFar;
    before_free = -kFar;
  }
  if (lane == kWarp - 1)
  {
    after_filled = kFar;
    after_free = kFar;
  }

  // 3. per voxel: nearest opposite-class voxel inside the word (bit scan) or outside (tables).
#pragma unroll
  for (int it = 0; it < kIterations; it++)
  {
    const int vector_index = (it << 5) + lane;
    const int w = (it << 2) + group;
    const int word_start = w << 5;
    // positions relative to the start of the word (all lanes shuffle; stores are predicated)
    const int left_of_filled = __shfl_sync(kFull, before_free, w) - word_start;
    const int left_of_free = __shfl_sync(kFull, before_filled, w) - word_start;
    const int right_of_filled = __shfl_sync(kFull, after_free, w) - word_start;
    const int right_of_free = __shfl_sync(kFull, after_filled, w) - word_start;
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
