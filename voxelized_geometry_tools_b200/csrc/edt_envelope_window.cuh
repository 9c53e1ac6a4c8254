// The strided-axis WINDOW kernel (sm_100a): the common case of the y / x passes without a stack.
// Included by edt_kernels.cu after edt_envelope_lean.cuh.
//
// The 1-D transform of a line is D(q) = min over rows i of G(i) + (q - i)^2, where G(i) is the
// partial squared distance of row i when it has the class of row q and 0 when it has the other
// class (sdfgen.cpp:85-226 computes exactly this for each of the reference's two fields). Two
// observations make most of it a fixed, branch-free amount of work:
//   * the class-agnostic minimum T(q) = min_i g(i) + (q - i)^2 over ALL rows differs from D(q)
//     only by the opposite-class rows, and those enter D(q) as (q - i)^2, whose minimum is e^2
//     with e the distance to the NEAREST opposite-class row: D(q) = min(T(q), e^2);
//   * a row further than R from q contributes at least (R + 1)^2. So the minimum W(q) over the
//     rows within R of q is the exact D(q) whenever W(q) < (R + 1)^2 - it certifies itself.
// Here one lane owns one line (a warp = 32 adjacent lines, every row access is one 128-byte
// segment, as in the stack kernels) and keeps the 3 R rows around the current chunk of R rows in
// registers, two rows per register as 16-bit halves, so T costs one VIADDMNMX.U16x2 per two
// candidate rows (R + 1 per voxel, the squared offsets are immediates). e comes from the
// boundary bits of a class-bit window: per row and side one shift and one count of leading
// zeros, then one three-way minimum and a look-up of the square across the warp (SHFL). The rows
// ahead arrive by plain loads one chunk ahead. No scratch, the input is read once and not
// modified: 4 B in + 4 B out per voxel.
//
// The kernel is bound by the integer alu pipe and by issue slots, not by HBM, so everything
// around the add-mins is written for the OTHER pipes: row addresses, shifts by constants, bit
// extraction and the doubling that drops the class bit are integer multiplies (fma pipe) whose
// power-of-two factors come from constant memory - ptxas would turn a literal back into an alu
// shift or LEA -, the bit scans run on the xu pipe, the table of squares is a warp shuffle, and
// the finalizing pass reads magnitudes of certified chunks from a copy of the head of the table
// in shared memory (one 32-bit address instead of a 64-bit one).
//
// Rows that do not certify (distance to the opposite class > R voxels) continue the same search
// outwards, four row pairs per vote, until d^2 reaches the best value so far - exact for any
// input but O(distance) per voxel, so the warp counts those steps and, past its allowance (or at
// a row in open space), hands its tile to the stack kernel (EnvelopeAxisLeanKernel over the redo
// list) instead. Distance fields of cluttered maps (distances of a few voxels nearly everywhere)
// stay on the fast path; for maps of open space a pilot launch (a probe over a sample of rows)
// decides on the device that the stack kernel takes every tile and this kernel stands down.
//
// Replaces the same reference loops as the stack kernels: the X / Y loops of
// ComputeDistanceFieldTransformInPlace (sdfgen.cpp:276-351) with the 1-D transforms
// (sdfgen.cpp:85-226) for both fields at once; in finalize mode also the combine loop
// (sdfgen.hpp:85-108) and Lock()'s min/max (sdf.hpp:765-787).
#pragma once

#include "edt_device.cuh"
#include "edt_envelope_lean.cuh"

namespace vgt_b200
{
namespace edt
{
namespace
{
// One warp (one tile) per block: warps finish at very different times (a tile through a deep
// pocket searches far longer than its neighbours), and a block's slot is only free again when
// its slowest warp is done. With 4-warp blocks the SMs held 25 warps on average out of the 32
// the registers allow.
constexpr int kWindowWarpsPerBlock = 1;
// Window radius and resident warps per SM (72 / 64 registers) for the packed (y) pass and the
// finalizing (x) pass: the in-plane distances the y pass sees are larger than the final ones.
#ifndef VGT_WINDOW_RADIUS_PACKED
#define VGT_WINDOW_RADIUS_PACKED 12
#endif
#ifndef VGT_WINDOW_RADIUS_FINAL
#define VGT_WINDOW_RADIUS_FINAL 8
#endif
constexpr int kWindowRadiusPacked = VGT_WINDOW_RADIUS_PACKED;
#ifndef VGT_WINDOW_BLOCKS_PACKED
#define VGT_WINDOW_BLOCKS_PACKED 28
#endif
constexpr int kWindowBlocksPacked = VGT_WINDOW_BLOCKS_PACKED;
constexpr int kWindowRadiusFinal = VGT_WINDOW_RADIUS_FINAL;
#ifndef VGT_WINDOW_BLOCKS_FINAL
#define VGT_WINDOW_BLOCKS_FINAL 32
#endif
constexpr int kWindowBlocksFinal = VGT_WINDOW_BLOCKS_FINAL;
// Hand-over buffer between the window kernel and the stack kernel ("redo list"), T = number of
// tiles: word 0 = number of listed tiles, word 1 = mode (0: the stack kernel redoes the listed
// tiles; 1: the pilot found the map deep, the stack kernel does every tile), word 2 = extended-
// search steps of the pilot's warps, word 3 unused, words 4 .. 4 + T = tile indices, then T
// per-tile "already listed" flags. The launcher zeroes everything but the list.
constexpr uint32_t kRedoCount = 0;
constexpr uint32_t kRedoMode = 1;
constexpr uint32_t kRedoPilotSteps = 2;
constexpr uint32_t kRedoList = 4;
// What a launch is: the plain pass over every tile; the pilot, a probe that computes two chunks
// of rows out of every kPilotSpacing rows of every kPilotStride-th tile, stores
// nothing and only reports its search effort (words 0 and 2), from which the mode is decided;
// or the pass after the pilot, which stands down when the mode is "deep".
constexpr uint32_t kSelectAll = 0;
constexpr uint32_t kSelectPilot = 1;
constexpr uint32_t kSelectAfterPilot = 2;
constexpr uint32_t kPilotStride = 8;
constexpr int kPilotSpacing = 256;
// Values in the register window are clamped to this (16-bit halves; the squared offsets of the
// joint search must still fit on top of it, see kJointDeepestCap).
constexpr uint32_t kSaturated = 0x7800u;

// Staged variant (kStage): the rows of the chunks ahead travel global -> shared memory with
// cp.async (4 bytes per lane and row, each lane later reads back only what it copied itself, so
// no barrier is needed), two chunks ahead, tracked by cp.async groups instead of the load
// scoreboards the register prefetch shares between its 12 loads in flight.
__device__ __forceinline__ void CopyRowAsync(uint32_t* shared_word, const void* global_word)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(shared_word))),
               "l"(global_word)
               : "memory");
}
__device__ __forceinline__ void CommitAsyncCopies()
{
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int kPending>
__device__ __forceinline__ void WaitForAsyncCopies()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

// Bulk variant (kStage == 2): the same three chunk buffers, filled by the TMA engine. One
// cp.async.bulk per row (128 contiguous bytes: the 32 columns of the tile), issued by lane j for
// row j - ONE predicated instruction per chunk instead of R per-lane copies with their address
// arithmetic - and one mbarrier per buffer armed with the bytes to expect; the warp waits on the
// barrier's phase parity where the cp.async variant waits for a copy group. Needs rows that start
// 16-byte aligned and full tiles (the launcher checks and falls back to kStage == 1).
__device__ __forceinline__ uint32_t SharedAddress(const void* pointer)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(pointer));
}
__device__ __forceinline__ void BarrierInit(uint64_t* barrier, uint32_t arrivals)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SharedAddress(barrier)),
               "r"(arrivals)
               : "memory");
}
__device__ __forceinline__ void BarrierArriveExpectBytes(uint64_t* barrier, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   SharedAddress(barrier)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void BarrierWait(uint64_t* barrier, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred done;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1;\n"
      "@!done bra WAIT_LOOP;\n"
      "}\n" ::"r"(SharedAddress(barrier)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void BulkCopyRow(uint32_t* shared_row, const void* global_row,
                                            uint32_t bytes, uint64_t* barrier)
{
  asm volatile(
      "cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          SharedAddress(shared_row)),
      "l"(global_row), "r"(bytes), "r"(SharedAddress(barrier))
      : "memory");
}

// Deepest row distance the 16-bit joint search looks at: clamped values (kSaturated) plus squared
// offsets must fit 16 bits and a result is only exact below kSaturated (0x7800: 175 voxels).
// The furthest candidate of a round is kJointDeepestCap + R - 1 rows from its row:
// (168 + 13)^2 + 0x7800 = 63481 < 65536.
constexpr int kJointDeepestCap = 168;
static_assert((kJointDeepestCap + 13) * (kJointDeepestCap + 13) + static_cast<int>(kSaturated) < 65536,
              "the joint search must stay inside 16 bits");
// "No opposite-class row in the neighbouring chunk" for the joint search: far enough that its
// square beats no real candidate, small enough that (kNoRow + R)^2 fits 16 bits.
constexpr uint32_t kNoRow = 200u;

// The window kernel is bound by the alu pipe (VIADDMNMX, LOP3, SHF, PRMT, VIMNMX, IADD3, LEA);
// the fma pipe next to it is nearly idle. Address arithmetic, shifts by constants and bit
// extraction can all be written as integer multiplies, which run on the fma pipe - but ptxas
// turns a multiply by a constant power of two back into a shift or LEA. So the multipliers come
// from constant memory, which ptxas cannot fold and which an IMAD takes as a direct operand.
__constant__ uint32_t kSmallNumbers[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
__constant__ uint32_t kPowersOfTwo[32] = {
    1u << 0,  1u << 1,  1u << 2,  1u << 3,  1u << 4,  1u << 5,  1u << 6,  1u << 7,
    1u << 8,  1u << 9,  1u << 10, 1u << 11, 1u << 12, 1u << 13, 1u << 14, 1u << 15,
    1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 21, 1u << 22, 1u << 23,
    1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30, 1u << 31};

// base + stride * row as ONE wide multiply-add instead of a 64-bit add chain.
template <typename Pointer>
__device__ __forceinline__ Pointer OffsetRows(Pointer base, uint32_t stride_bytes, uint32_t row)
{
  uint64_t address;
  asm("mad.wide.u32 %0, %1, %2, %3;"
      : "=l"(address)
      : "r"(stride_bytes), "r"(row), "l"(reinterpret_cast<uint64_t>(base)));
  return reinterpret_cast<Pointer>(address);
}
// (row known at compile time, 0 .. 15)
template <int kRow, typename Pointer>
__device__ __forceinline__ Pointer OffsetRowsBy(Pointer base, uint32_t stride_bytes)
{
  static_assert(kRow >= 0 && kRow < 16, "kSmallNumbers");
  return OffsetRows(base, stride_bytes, kSmallNumbers[kRow]);
}
// word << kBits and word >> kBits on the fma pipe
template <int kBits>
__device__ __forceinline__ uint32_t ShiftLeftByMultiply(uint32_t word)
{
  static_assert(kBits >= 0 && kBits < 32, "kPowersOfTwo");
  return word * kPowersOfTwo[kBits];
}
template <int kBits>
__device__ __forceinline__ uint32_t ShiftRightByMultiply(uint32_t word)
{
  static_assert(kBits >= 1 && kBits <= 32, "kPowersOfTwo");
  return __umulhi(word, kPowersOfTwo[32 - kBits]);
}

// A row word of the input. The addresses come out of 64-bit arithmetic the compiler no longer
// knows the address space of; said explicitly, the load is an LDG instead of a generic LD.
#ifndef VGT_WINDOW_GLOBAL_LOADS
#define VGT_WINDOW_GLOBAL_LOADS 1
#endif
__device__ __forceinline__ uint32_t LoadRowWord(const void* address)
{
#if VGT_WINDOW_GLOBAL_LOADS == 1
  uint32_t word;
  asm("ld.global.u32 %0, [%1];" : "=r"(word) : "l"(address));
  return word;
#elif VGT_WINDOW_GLOBAL_LOADS == 2
  return __ldg(static_cast<const uint32_t*>(address));
#else
  return *static_cast<const uint32_t*>(address);
#endif
}

// Number of leading zero bits; 0xffffffff when the word is 0 (one FLO).
__device__ __forceinline__ uint32_t LeadingZeros(uint32_t word)
{
  uint32_t count;
  asm("bfind.shiftamt.u32 %0, %1;" : "=r"(count) : "r"(word));
  return count;
}

// Build-time experiment switches (profiles/r2_experiments.md records the outcomes).
#ifndef VGT_WINDOW_ONE_CLASS_PATH
#define VGT_WINDOW_ONE_CLASS_PATH 0
#endif
#ifndef VGT_WINDOW_SHIFT_CHAINS
#define VGT_WINDOW_SHIFT_CHAINS 1
#endif

// Accumulators per row of the window minimum (the R rows of a chunk are independent chains
// already; more than one accumulator per row costs a merge per row).
#ifndef VGT_WINDOW_CHAINS
#define VGT_WINDOW_CHAINS 1
#endif
constexpr int kWindowChains = VGT_WINDOW_CHAINS;

// Calls f(std::integral_constant<int, 0>{}) ... f(std::integral_constant<int, kCount - 1>{}).
template <int kCount, int kIndex = 0, typename F>
__device__ __forceinline__ void StaticFor(F&& f)
{
  if constexpr (kIndex < kCount)
  {
    f(std::integral_constant<int, kIndex>{});
    StaticFor<kCount, kIndex + 1>(f);
  }
}

// min(best, (c + a_low)^2 | (c + a_high)^2 << 16) per 16-bit half; c_squared = c * c * 0x10001.
template <int kLow, int kHigh>
__device__ __forceinline__ uint32_t FoldSquaredDistances(uint32_t best, uint32_t c,
                                                         uint32_t c_squared)
{
  constexpr uint32_t kLinear = static_cast<uint32_t>(2 * kLow) | (static_cast<uint32_t>(2 * kHigh) << 16);
  constexpr uint32_t kConstant =
      static_cast<uint32_t>(kLow * kLow) | (static_cast<uint32_t>(kHigh * kHigh) << 16);
  return __viaddmin_u16x2(c * kLinear + c_squared, kConstant, best);
}

// Work split: blockIdx.x = tile (the pilot launch: tile blockIdx.x * kPilotStride),
// blockIdx.y = segment: rows [blockIdx.y * segment_spacing, + segment_rows) of their lines
// (segment_rows a multiple of R; spacing == rows except for the pilot's probes).

// After the pilot, whose probes run on a tight allowance (kPilotStepRate: two search steps per
// row, so that a probe in open space gives up at its first or second deep row instead of
// searching for 100 microseconds): the map is "deep" when 40 % of the probes gave up. Then the
// window launch stands down and the stack kernel takes every tile. Resets the count for the real
// pass.
constexpr uint32_t kPilotStepRate = 2u * 128u;
__global__ void DecideWindowModeKernel(uint32_t* redo, uint32_t pilot_probes)
{
  const bool deep = 5u * redo[kRedoCount] > 2u * pilot_probes;
  redo[kRedoMode] = deep ? 1u : 0u;
  redo[kRedoCount] = 0u;
}

template <int kMode, int kR, bool kBorder, bool kSend, int kBlocksPerSm, int kStage = 0>
__global__ void __launch_bounds__(kWindowWarpsPerBlock* kWarp, kBlocksPerSm)
    EnvelopeAxisWindowKernel(const uint32_t* __restrict__ in,
                             typename OutputOf<kMode>::Type* __restrict__ out, LineFamily family,
                             FinalizeParams finalize, typename OutputOf<kMode>::Key* min_max_keys,
                             uint32_t* redo, uint32_t step_rate, int segment_rows,
                             int segment_spacing, uint32_t block_select)
{
  using Out = typename OutputOf<kMode>::Type;
  static_assert(kR >= 2 && kR <= 14 && kR % 2 == 0,
                "the class window is one 32-bit word (2 R + 1 <= 31 bits); rows are kept in pairs");
  static_assert((kR + 1) * (kR + 1) <= static_cast<int>(kSaturated), "kSaturated >= kFar");
  constexpr uint32_t kFar = static_cast<uint32_t>((kR + 1) * (kR + 1));
  constexpr uint32_t kSideMask = (1u << kR) - 1u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int length = family.length;
  const int last_row = static_cast<int>(family.last_row);

  const uint32_t tiles_per_outer = static_cast<uint32_t>((family.inner_count + kWarp - 1) / kWarp);
  const uint32_t block_index = (block_select == kSelectPilot) ? blockIdx.x * kPilotStride : blockIdx.x;
  if (block_select == kSelectAfterPilot && redo[kRedoMode] != 0u)
  {
    return;  // a deep map: the stack kernel does it all
  }
  const uint32_t tile_index = block_index * kWindowWarpsPerBlock + warp;
  if (tile_index >= tiles_per_outer * static_cast<uint32_t>(family.num_outer))
  {
    return;  // warp-uniform
  }
  const uint32_t outer = tile_index / tiles_per_outer;
  const uint32_t wanted_column = (tile_index - outer * tiles_per_outer) * kWarp + lane;
  // Lanes past the last column shadow the last one, so that the whole warp stays converged for
  // the votes: they compute and store exactly what the lane of that column stores, to the same
  // addresses (no predicate on the stores). The pilot stores nothing - it skips phase C - and
  // all its lanes read the family's last column (one sector per row).
  const bool probe_only = block_select == kSelectPilot;
  const uint32_t column = probe_only
      ? static_cast<uint32_t>(family.inner_count - 1)
      : min(wanted_column, static_cast<uint32_t>(family.inner_count - 1));
  const int64_t first = static_cast<int64_t>(outer) * family.outer_stride + column;
  // The two strides live in per-lane registers that ptxas cannot prove uniform (the top bit of
  // the lane's column is 0: a family has fewer than 2^31 columns): otherwise it forms
  // stride * row in the uniform datapath and adds the product to the lane's pointer with a
  // two-instruction 64-bit add on the alu pipe; this way a row address is one IMAD.WIDE on the
  // fma pipe.
  const uint32_t lane_zero = column >> 31;
  const uint32_t stride_bytes = family.stride_bytes + lane_zero;
  // (pinned in registers: the compiler otherwise re-derives both from the kernel parameters at
  // every use inside the unrolled chunk bodies)
  const char* line = reinterpret_cast<const char*>(in + first);
  asm volatile("" : "+l"(line));
  const auto load_row = [&](int row)
  {
    return LoadRowWord(line + static_cast<uint64_t>(static_cast<uint32_t>(row)) * stride_bytes);
  };
  char* write_origin = reinterpret_cast<char*>(out + first);
  asm volatile("" : "+l"(write_origin));
  const uint32_t out_stride_bytes = family.out_stride_bytes + lane_zero;

  // Finalize mode: the largest squared distance this lane has emitted for a free voxel (low
  // half) and for a filled voxel (high half). The magnitude is monotone in the squared distance,
  // so these two are what the lane contributes to Lock()'s maximum and minimum; every emitted
  // value is finite and below kSaturated (anything else gives the tile up), hence 16 bits, and a
  // finite value means both classes exist, so the smallest free / filled values cannot be the
  // extrema. 0 = no voxel of that class emitted (a real squared distance is at least 1).
  uint32_t lane_extrema = 0;
  // Float output: the extrema straight from the bits of the emitted values. As signed integers
  // every positive float is above every negative one and positive floats order like their bits:
  // the signed maximum is the largest free value (still negative: no free voxel emitted). As
  // unsigned integers negative floats are above positive ones and order by magnitude: the
  // unsigned maximum is the most negative filled value (sign bit clear: no filled voxel).
  int32_t largest_signed = static_cast<int32_t>(0x80000000u);
  uint32_t largest_unsigned = 0u;
  // Finalize mode: the head of the magnitude table (squares below kFar, all a certified chunk
  // can emit) copied into shared memory once per block.
  __shared__ Out near_magnitudes[(kMode != kEmitPacked) ? kFar : 1];
  if constexpr (kMode != kEmitPacked)
  {
    static_assert(kWindowWarpsPerBlock == 1, "one copy per warp: only __syncwarp orders it");
    const Out* const table = static_cast<const Out*>(finalize.magnitude_table);
    for (uint32_t i = lane; i < kFar; i += kWarp)
    {
      near_magnitudes[i] = __ldg(table + i);
    }
    __syncwarp();
  }
  // The squared distance to the nearest opposite-class row as a look-up across the warp: lane d - 1
  // holds d^2 for d = 1 .. R and 0xffff ("no such row") otherwise, once in the low half (even
  // rows) and once in the high half (odd rows), the other half 0xffff, so that one three-way
  // paired minimum merges both rows of a pair.
  const uint32_t lane_square = (lane < kR) ? static_cast<uint32_t>((lane + 1) * (lane + 1)) : 0xffffu;
  uint32_t squares_low_half = 0xffff0000u | lane_square;
  uint32_t squares_high_half = (lane_square << 16) | 0xffffu;
  asm volatile("" : "+r"(squares_low_half), "+r"(squares_high_half));
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

  // The three chunks of R rows around the rows being computed, two rows per register: 16-bit
  // halves, low half = the even row of the pair, so that one VIADDMNMX.U16x2 handles two
  // candidates. Values are clamped to kSaturated: a candidate that large cannot certify anything
  // (kSaturated >= kFar), and as long as the window minimum stays below kSaturated it was taken
  // over unclamped values only, i.e. it is exact. Rows outside the line hold kSaturated and
  // repeat the class of the nearest row of the line, so they never read as the nearest
  // opposite-class row.
  constexpr int kPairs = kR / 2;
  uint32_t previous_pairs[kPairs], current_pairs[kPairs], next_pairs[kPairs];
  // Class bits of the three chunks: bit 3 R - 1 - r' = class of row r' counted from the first
  // row of the previous chunk (once the next chunk has been absorbed at the start of a chunk:
  // every absorbed row shifts its bit in at the bottom); bits above 3 R are leftovers of older
  // chunks and never reach a window.
  // (one 32-bit register when the three chunks fit, i.e. R <= 10: every class operation is then
  // a single instruction instead of a 64-bit pair)
  using ClassWord = typename std::conditional<(3 * kR <= 32), uint32_t, uint64_t>::type;
  ClassWord classes = 0;
  // The rows of the next chunk as loaded. Rows j, j + 1 (j even) of it are needed from row j of
  // the current chunk on, so they are absorbed (clamped, packed, class bits filed) right there,
  // and each register is reloaded with the row one chunk further after its own row: every load
  // has about a chunk of work to land.
  uint32_t raw[kStage ? 1 : kR];
  // kStage: three chunk buffers of R rows x 32 lanes in shared memory instead (one warp per
  // block); absorb_buffer holds the next chunk, the one after it is in flight, fill_buffer takes
  // the chunk three ahead.
  __shared__ alignas(128) uint32_t stage[kStage ? 3 * kR * kWarp : 1];
  static_assert(!kStage || kWindowWarpsPerBlock == 1, "the stage buffers are per block");
  uint32_t* const stage_lane = stage + lane;
  int absorb_buffer = 0;
  int fill_buffer = 2;
  // kStage == 2: one mbarrier per buffer; bit b of wait_parity = the phase parity the next wait
  // on buffer b looks for, bit b of in_flight = a fill of buffer b has not been waited for yet.
  __shared__ alignas(8) uint64_t stage_barriers[kStage == 2 ? 3 : 1];
  uint32_t wait_parity = 0;
  uint32_t in_flight = 0;
  if constexpr (kStage == 2)
  {
    if (lane == 0)
    {
#pragma unroll
      for (int b = 0; b < 3; b++)
      {
        BarrierInit(stage_barriers + b, 1u);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }

  const auto load_clamped = [&](const int row) { return load_row(min(max(row, 0), last_row)); };
  const auto clamped_value = [&](const uint32_t word) { return min(word & kNone, kSaturated); };

  // kStage: the R rows starting at `first` (kClamp: clamped to the line) go into stage buffer
  // `buffer`; and the wait for the rows of a buffer to have landed.
  const auto fill_stage = [&](const int buffer, const int first, auto clamp)
  {
    constexpr bool kClamp = decltype(clamp)::value;
    if constexpr (kStage == 1)
    {
#pragma unroll
      for (int i = 0; i < kR; i++)
      {
        const int row = kClamp ? min(first + i, last_row) : first + i;
        CopyRowAsync(stage_lane + (buffer * kR + i) * kWarp,
                     line + static_cast<uint64_t>(static_cast<uint32_t>(row)) * stride_bytes);
      }
      CommitAsyncCopies();
    }
    else if constexpr (kStage == 2)
    {
      if (lane == 0)
      {
        BarrierArriveExpectBytes(stage_barriers + buffer, kR * kWarp * 4u);
      }
      if (lane < kR)
      {
        // lane i issues row i: 128 bytes from the tile's first column (line is this lane's own
        // column; full tiles only, so the tile starts lane columns before it)
        const int row = kClamp ? min(first + lane, last_row) : first + lane;
        BulkCopyRow(stage + (buffer * kR + lane) * kWarp,
                    line - 4 * lane + static_cast<uint64_t>(static_cast<uint32_t>(row)) * stride_bytes,
                    kWarp * 4u, stage_barriers + buffer);
      }
      in_flight |= 1u << buffer;
    }
  };
  const auto wait_for_stage = [&](const int buffer)
  {
    if constexpr (kStage == 1)
    {
      WaitForAsyncCopies<1>();  // everything but the newest group
    }
    else if constexpr (kStage == 2)
    {
      BarrierWait(stage_barriers + buffer, (wait_parity >> buffer) & 1u);
      wait_parity ^= 1u << buffer;
      in_flight &= ~(1u << buffer);
    }
  };

  // Extended-search steps of this warp so far (warp-uniform) and whether they have passed the
  // allowance of step_rate / 128 steps per row done (plus a credit of a quarter segment), or a
  // row turned out deeper than deepest_search: the tile then goes to the stack kernel.
  uint32_t steps = 0;
  bool over_budget = false;
  // Fused exchange: the segments are taken in an order rotated by the rank and by the line's
  // outer index, so that the blocks in flight at any moment cover all parts of the lines and
  // every rank stores into all its peers at once (see LineFamily::scatter_rank).
  uint32_t segment = blockIdx.y;
  if constexpr (kSend)
  {
    if (family.scatter_base[0] != nullptr)
    {
      segment = (segment + family.first_segment + outer) % gridDim.y;
    }
  }
  const int first_row = static_cast<int>(segment) * segment_spacing;
  const int end_row = min(first_row + segment_rows, length);
  if (first_row >= length)
  {
    return;  // warp-uniform
  }
  const uint32_t credit_rows = static_cast<uint32_t>(segment_rows >> 2) + 16u;
  const int deepest_search = kJointDeepestCap;

  // Send layout only: the virtual origin of the current part and the row at which the next
  // part starts (the segment's first row forces the first look-up).
  char* send_origin = nullptr;
  int next_part_start = first_row;

  // Send layout (see LineFamily): row q of a line is at send_origin + q * out_stride_bytes; the
  // origin jumps at part boundaries, which are the same for every lane. locate_part(q) finds
  // the part that holds row q: its origin and the first row of the next part.
  const auto locate_part = [&](const int q)
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
      uint32_t* target = family.scatter_base[0];
#pragma unroll
      for (int i = 1; i < 8; i++)
      {
        target = (part == i) ? family.scatter_base[i] : target;
      }
      part_row = reinterpret_cast<Out*>(target)
          + family.inner_count * ((family.scatter_row_offset + outer) * rows + (q - y0))
          + column;
    }
    else
    {
      part_row = out + family.inner_count * (family.num_outer * y0 + outer * rows + (q - y0))
          + column;
    }
    send_origin = reinterpret_cast<char*>(part_row)
        - static_cast<uint64_t>(static_cast<uint32_t>(q)) * out_stride_bytes;
  };

  // The packed words of a chunk (words[j] = row base + j; `rows` of them are inside the line).
  // Plain layout: row j at write_base + j * stride. Send layout: the rows are emitted group by
  // group, a group = the rows of the chunk that lie in one part (nearly always the whole
  // chunk), with ONE part look-up per group: the look-up (two divisions) exists once in the
  // code, not once per unrolled row.
  const auto emit_packed = [&](const int base, char* const write_base, const uint32_t* words,
                               const int rows)
  {
    const auto store = [&](char* const at, const uint32_t word)
    {
      // (peer stores: .cs / .cg / .wt / default measured the same, profiles/r2_experiments.md)
      __stcs(reinterpret_cast<uint32_t*>(at), word);
    };
    if constexpr (!kSend)
    {
#pragma unroll
      for (int j = 0; j < kR; j++)
      {
        if (j < rows)  // warp-uniform (always true in interior chunks)
        {
          store(OffsetRows(write_base, out_stride_bytes, kSmallNumbers[j]), words[j]);
        }
      }
    }
    else
    {
      int row = 0;
#pragma unroll 1
      while (row < rows)
      {
        if (base + row >= next_part_start)  // warp-uniform
        {
          locate_part(base + row);
        }
        const int group_end = min(rows, next_part_start - base);
        char* const group_base = OffsetRows(send_origin, out_stride_bytes, static_cast<uint32_t>(base));
#pragma unroll
        for (int j = 0; j < kR; j++)
        {
          if (j >= row && j < group_end)  // warp-uniform
          {
            store(OffsetRows(group_base, out_stride_bytes, kSmallNumbers[j]), words[j]);
          }
        }
        row = group_end;
      }
    }
  };

  // Finalize mode, a whole chunk: the virtual border, R magnitude look-ups in flight together
  // (the per-call table holds (Out)(sqrt((double)s) * resolution) for every s below kSaturated at
  // least, see MagnitudeTable::Build: every value this kernel emits is inside it), then sign,
  // store and the two integer extrema.
  const auto emit_finalized = [&](const int base, char* const write_base, const uint32_t* best,
                                  auto edge, auto near)
  {
    constexpr bool kEdge = decltype(edge)::value;
    // kNear: every row of the chunk certified itself (all squares below kFar): the magnitudes
    // come from the block's copy of the head of the table in shared memory - a 32-bit address,
    // one multiply-add - instead of a 64-bit address into the global table
    constexpr bool kNear = decltype(near)::value;
    // (a per-lane pointer, see lane_zero: the look-up address is one IMAD.WIDE)
    const Out* const table = static_cast<const Out*>(finalize.magnitude_table) + lane_zero;
    Out magnitudes[kR];
    uint32_t squares[kR];
#pragma unroll
    for (int j = 0; j < kR; j++)
    {
      const int q = base + j;
      // (the odd row of a pair comes down by a multiply: fma pipe)
      uint32_t squared = (j & 1) ? ShiftRightByMultiply<16>(best[j >> 1]) : (best[j >> 1] & 0xffffu);
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
      squares[j] = squared;
      magnitudes[j] = Out(0);
      if (!kEdge || q <= last_row)  // warp-uniform
      {
        if constexpr (kNear)
        {
          magnitudes[j] = *reinterpret_cast<const Out*>(
              reinterpret_cast<const char*>(near_magnitudes) + squared * kSmallNumbers[sizeof(Out)]);
        }
        else
        {
          magnitudes[j] = __ldg(OffsetRows(table, kSmallNumbers[sizeof(Out)], squared));
        }
      }
    }
    if constexpr (sizeof(Out) == 4)
    {
      uint32_t bits[kR];
#pragma unroll
      for (int j = 0; j < kR; j++)
      {
        // the class of row j in bit 31, flipped into the sign of the magnitude (never zero)
        const uint32_t sign = static_cast<uint32_t>(classes) * kPowersOfTwo[32 - 2 * kR + j];
        bits[j] = __float_as_uint(magnitudes[j]) ^ (sign & 0x80000000u);
        if (!kEdge || base + j <= last_row)  // warp-uniform
        {
          __stcs(reinterpret_cast<uint32_t*>(
                     OffsetRows(write_base, out_stride_bytes, kSmallNumbers[j])),
                 bits[j]);
        }
        else
        {
          bits[j] = bits[0];  // (row 0 of a chunk is always inside the line)
        }
      }
#pragma unroll
      for (int j = 0; j < kR; j += 2)
      {
        largest_signed = __vimax3_s32(largest_signed, static_cast<int32_t>(bits[j]),
                                      static_cast<int32_t>(bits[j + 1]));
        largest_unsigned = __vimax3_u32(largest_unsigned, bits[j], bits[j + 1]);
      }
    }
    else
    {
#pragma unroll
      for (int j = 0; j < kR; j++)
      {
        const int q = base + j;
        if (!kEdge || q <= last_row)  // warp-uniform
        {
          const uint32_t filled = static_cast<uint32_t>(classes >> (2 * kR - 1 - j)) & 1u;
          const Out value = filled ? -magnitudes[j] : magnitudes[j];
          __stcs(reinterpret_cast<Out*>(
                     OffsetRows(write_base, out_stride_bytes, kSmallNumbers[j])),
                 value);
          // free: the low half, filled: the high half (one multiply on the fma pipe)
          lane_extrema = __vmaxu2(lane_extrema, squares[j] * (filled * 0xffffu + 1u));
        }
      }
    }
  };

  // One chunk: the rows base .. base + R - 1 (kEdge: only those inside the line), and the loads
  // of the rows base + 2 R .. base + 3 R - 1 (kEdge: clamped to the line) into `raw`.
  //   phase A  every row's window minimum and nearest opposite-class row inside the window,
  //            results as 16-bit pairs (low half = the even row);
  //   phase B  only when some lane holds a row that did not certify itself: first the rows of
  //            the register window beyond each row's own window, then rows from memory, each
  //            of which serves ALL rows of the chunk (one load, R / 2 paired add-mins);
  //   phase C  the rows are emitted.
  const auto compute_chunk = [&](const int base, auto edge)
  {
    constexpr bool kEdge = decltype(edge)::value;
    // (kStage: the chunk three ahead is fetched, else the chunk two ahead)
    constexpr int kAhead = kStage ? 3 : 2;
    const char* const read_next =
        line + static_cast<uint64_t>(stride_bytes) * static_cast<uint32_t>(base + kAhead * kR);
    if constexpr (kStage != 0)
    {
      wait_for_stage(absorb_buffer);  // the next chunk has landed
    }
    char* const write_base =
        write_origin + static_cast<uint64_t>(out_stride_bytes) * static_cast<uint32_t>(base);
    uint32_t best_pairs[kPairs];
    // ---------------------------------------------------------------------------------- phase A
    // The next chunk (rows base + R .. base + 2 R - 1) joins the register window: clamped, packed
    // in pairs, class bits filed. Then the loads of the chunk after it are issued: they have
    // the whole chunk to land.
    // (multiplies instead of masks, shifts and funnel shifts: they run on the fma pipe. Twice a
    // word drops its class bit; the doubled values are clamped, packed and halved together.)
    uint32_t chunk_classes = 0;
#pragma unroll
    for (int j = 0; j < kR; j += 2)
    {
      const uint32_t low_word =
          kStage ? stage_lane[(absorb_buffer * kR + j) * kWarp] : raw[kStage ? 0 : j];
      const uint32_t high_word =
          kStage ? stage_lane[(absorb_buffer * kR + j + 1) * kWarp] : raw[kStage ? 0 : j + 1];
      // (one wide multiply by two: the low word is twice the value, the high word the class bit)
      const uint64_t low_wide = static_cast<uint64_t>(low_word) * kPowersOfTwo[1];
      const uint64_t high_wide = static_cast<uint64_t>(high_word) * kPowersOfTwo[1];
      uint32_t low = min(static_cast<uint32_t>(low_wide), 2u * kSaturated);
      uint32_t high = min(static_cast<uint32_t>(high_wide), 2u * kSaturated);
      if constexpr (kEdge)
      {
        low = (base + j + kR > last_row) ? 2u * kSaturated : low;
        high = (base + j + kR + 1 > last_row) ? 2u * kSaturated : high;
      }
      next_pairs[j >> 1] = ShiftRightByMultiply<1>(__byte_perm(low, high, 0x5410));
      chunk_classes = chunk_classes * kPowersOfTwo[1] + static_cast<uint32_t>(low_wide >> 32);
      chunk_classes = chunk_classes * kPowersOfTwo[1] + static_cast<uint32_t>(high_wide >> 32);
    }
    if constexpr (sizeof(ClassWord) == 4)
    {
      classes = classes * kPowersOfTwo[kR] + chunk_classes;
    }
    else
    {
      classes = (classes << kR) | chunk_classes;
    }
    if constexpr (kStage != 0)
    {
      // the chunk three ahead goes into the buffer whose rows were absorbed during the
      // previous chunk
      fill_stage(fill_buffer, base + kAhead * kR, std::integral_constant<bool, kEdge>{});
      fill_buffer = absorb_buffer;
      absorb_buffer = (absorb_buffer == 2) ? 0 : absorb_buffer + 1;
    }
    else
    {
#pragma unroll
      for (int j = 0; j < kR; j++)
      {
        const char* const next_row = kEdge
            ? OffsetRows(line, stride_bytes,
                         static_cast<uint32_t>(min(base + kAhead * kR + j, last_row)))
            : OffsetRows(read_next, stride_bytes, kSmallNumbers[j]);
        raw[j] = LoadRowWord(next_row);
      }
    }
    // Every row's window minimum, and (kWithClasses) the nearest opposite-class row inside its
    // window. Chunks in which every lane's three register chunks are of ONE class - the bulk of
    // free space and of the inside of obstacles - skip the class arithmetic altogether.
    const auto window_rows = [&](auto with_classes)
    {
      constexpr bool kWithClasses = decltype(with_classes)::value;
      // Boundary bits of the class word: bit t = the row at bit t and the (earlier) row at bit
      // t + 1 are of different classes. For the row at bit p, the nearest opposite-class row
      // AFTER it is the highest boundary bit below p, the nearest one BEFORE it the lowest
      // boundary bit at or above p, i.e. the highest bit of the reversed word below 32 - p. Each
      // side: one shift that drops the bits beyond the row (a multiply, fma pipe) and one
      // count of leading zeros = distance - 1 (xu pipe); then ONE three-way minimum with R
      // (= "none inside the window") and a look-up across the warp.
      // kShift: with a 64-bit class word (3 R > 32) the earlier side looks at bits kShift ..
      // kShift + 31.
      constexpr int kShift = (3 * kR > 32) ? 3 * kR - 32 : 0;
      uint32_t boundaries_low = 0;
      uint32_t boundaries_reversed = 0;
      if constexpr (kWithClasses)
      {
        const ClassWord boundaries = classes ^ (classes >> 1);
        boundaries_low = static_cast<uint32_t>(boundaries);
        boundaries_reversed = __brev(static_cast<uint32_t>(boundaries >> kShift));
      }
      // (VGT_WINDOW_SHIFT_CHAINS: the shifted boundary words of the R rows as two chains of
      // doublings - one constant instead of 2 R different ones)
      uint32_t after_words[kR];
      uint32_t before_words[kR];
      if constexpr (kWithClasses && VGT_WINDOW_SHIFT_CHAINS != 0)
      {
        // row j sits at bit p = 2 R - 1 - j: "after" shifts by 32 - p = 33 - 2 R + j (grows
        // with j), "before" by p - kShift (grows towards row 0)
        after_words[0] = boundaries_low * kPowersOfTwo[33 - 2 * kR];
        before_words[kR - 1] = boundaries_reversed * kPowersOfTwo[kR - kShift];
#pragma unroll
        for (int j = 1; j < kR; j++)
        {
          after_words[j] = after_words[j - 1] * kPowersOfTwo[1];
          before_words[kR - 1 - j] = before_words[kR - j] * kPowersOfTwo[1];
        }
      }
#pragma unroll
      for (int j = 0; j < kR; j += 2)
      {
        // rows j and j + 1 together: their two accumulators are transposed into (low halves,
        // high halves) so that ONE paired minimum finishes both and leaves them packed
        uint32_t row_pair[2];
        uint32_t nearest_pair[2];
#pragma unroll
        for (int r = 0; r < 2; r++)
        {
          const int jr = j + r;
          // Pairs g = (jr >> 1) .. (jr >> 1) + R of the 3 R / 2 pairs in registers cover the
          // rows q - R .. q + R plus one row at distance R + 1 (a true candidate like the
          // others).
          uint32_t chains[kWindowChains];
#pragma unroll
          for (int c = 0; c < kWindowChains; c++)
          {
            chains[c] = 0xffffffffu;
          }
#pragma unroll
          for (int t = 0; t <= kR; t++)
          {
            const int g = (jr >> 1) + t;                // pair index, 0 .. 3 R / 2 - 1
            const int low_row = 2 * g - kR;             // chunk-relative row of the low half
            const int d_low = (jr > low_row) ? jr - low_row : low_row - jr;
            const int d_high = (jr > low_row + 1) ? jr - low_row - 1 : low_row + 1 - jr;
            const uint32_t offsets = static_cast<uint32_t>(d_low * d_low)
                | (static_cast<uint32_t>(d_high * d_high) << 16);
            const uint32_t pair = (g < kPairs)
                ? previous_pairs[(g < kPairs) ? g : 0]
                : ((g < 2 * kPairs)
                       ? current_pairs[(g >= kPairs && g < 2 * kPairs) ? g - kPairs : 0]
                       : next_pairs[(g >= 2 * kPairs) ? g - 2 * kPairs : 0]);
            chains[t % kWindowChains] = __viaddmin_u16x2(pair, offsets, chains[t % kWindowChains]);
          }
          row_pair[r] = chains[0];
#pragma unroll
          for (int c = 1; c < kWindowChains; c++)
          {
            row_pair[r] = __vminu2(row_pair[r], chains[c]);
          }
          if constexpr (kWithClasses)
          {
            // the row sits at bit p of the class word
            const int p = 2 * kR - 1 - jr;
            const uint32_t after = LeadingZeros(
                VGT_WINDOW_SHIFT_CHAINS != 0 ? after_words[jr] : boundaries_low * kPowersOfTwo[32 - p]);
            const uint32_t before = LeadingZeros(
                VGT_WINDOW_SHIFT_CHAINS != 0 ? before_words[jr]
                                             : boundaries_reversed * kPowersOfTwo[p - kShift]);
            // squared, in the half of its row (0xffff in the other half and for "none")
            nearest_pair[r] = __shfl_sync(0xffffffffu, r == 0 ? squares_low_half : squares_high_half,
                                          __vimin3_u32(after, before, static_cast<uint32_t>(kR)));
          }
        }
        uint32_t both;
        if constexpr (kWithClasses)
        {
          both = __vimin3_u16x2(__byte_perm(row_pair[0], row_pair[1], 0x5410),
                                __byte_perm(row_pair[0], row_pair[1], 0x7632),
                                __vminu2(nearest_pair[0], nearest_pair[1]));
        }
        else
        {
          both = __vminu2(__byte_perm(row_pair[0], row_pair[1], 0x5410),
                          __byte_perm(row_pair[0], row_pair[1], 0x7632));
        }
        if constexpr (kEdge)
        {
          // (a row past the end of the line: 0, so that it never asks for a search)
          both = (base + j > last_row) ? 0u : ((base + j + 1 > last_row) ? (both & 0xffffu) : both);
        }
        best_pairs[j >> 1] = both;
      }
    };
    bool one_class_everywhere = false;
    if constexpr (!kEdge && VGT_WINDOW_ONE_CLASS_PATH != 0)
    {
      constexpr ClassWord kSpan = (static_cast<ClassWord>(1) << (3 * kR)) - 1;
      const ClassWord span = classes & kSpan;
      one_class_everywhere = __all_sync(0xffffffffu, span == 0 || span == kSpan);
    }
    if (one_class_everywhere)
    {
      window_rows(std::false_type{});
    }
    else
    {
      window_rows(std::true_type{});
    }
    // ---------------------------------------------------------------------------------- phase B
    // The largest result of this lane's rows: a row is certified iff its result is below kFar.
    const auto worst_of_lane = [&]()
    {
      uint32_t worst_pair = best_pairs[0];
#pragma unroll
      for (int p = 1; p < kPairs; p++)
      {
        worst_pair = __vmaxu2(worst_pair, best_pairs[p]);
      }
      return max(worst_pair & 0xffffu, worst_pair >> 16);
    };
    uint32_t worst = worst_of_lane();
    const bool searched = __any_sync(0xffffffffu, worst >= kFar);
    if (searched)
    {
      // A lane with an uncertified row has no opposite-class row within R of that row, so all
      // its rows of this chunk are of one class, the class of row `base`. (Lanes whose chunk is
      // of both classes hold certified rows only - bounded by (R - 1)^2 - and no candidate
      // below, at distance > R, can win against those.)
      const uint32_t chunk_bits = static_cast<uint32_t>(classes >> kR) & kSideMask;
      const uint32_t chunk_filled = (chunk_bits >> (kR - 1)) & 1u;
      const uint32_t flip = chunk_filled ? kSideMask : 0u;
      const bool one_class = chunk_bits == flip;
      // Nearest opposite-class row of the previous chunk (distance c from row `base`) and of
      // the next chunk (distance c from the chunk's last row): candidates (c + j)^2 and
      // (c + R - 1 - j)^2 for row j.
      {
        const uint32_t differs_before = (static_cast<uint32_t>(classes >> (2 * kR)) ^ flip) & kSideMask;
        const uint32_t differs_after = (static_cast<uint32_t>(classes) ^ flip) & kSideMask;
        const uint32_t below = (one_class && differs_before != 0u)
            ? static_cast<uint32_t>(__ffs(static_cast<int>(differs_before)))
            : kNoRow;
        const uint32_t above = (one_class && differs_after != 0u)
            ? static_cast<uint32_t>(kR - 31 + __clz(static_cast<int>(differs_after)))
            : kNoRow;
        const uint32_t below_squared = below * below * 0x10001u;
        const uint32_t above_squared = above * above * 0x10001u;
        const auto fold = [&](auto pair_index)
        {
          constexpr int p = decltype(pair_index)::value;
          best_pairs[p] = FoldSquaredDistances<2 * p, 2 * p + 1>(best_pairs[p], below, below_squared);
          best_pairs[p] = FoldSquaredDistances<kR - 1 - 2 * p, kR - 2 - 2 * p>(best_pairs[p], above,
                                                                             above_squared);
        };
        StaticFor<kPairs>(fold);
      }
      // Class-agnostic candidates: the rows of the register window a row's own window did not
      // reach. Row k of the previous chunk is missing for the rows j > k, at distance j + R - k;
      // row m of the next chunk for the rows j < m, at distance R + m - j.
#pragma unroll
      for (int k = 0; k < kR - 1; k++)
      {
        const uint32_t candidate = __byte_perm(previous_pairs[k >> 1], 0u, (k & 1) ? 0x3232 : 0x1010);
#pragma unroll
        for (int p = 0; p < kPairs; p++)
        {
          if (2 * p + 1 > k)
          {
            const int d_low = 2 * p + kR - k;
            const uint32_t offsets = static_cast<uint32_t>(d_low * d_low)
                | (static_cast<uint32_t>((d_low + 1) * (d_low + 1)) << 16);
            best_pairs[p] = __viaddmin_u16x2(candidate, offsets, best_pairs[p]);
          }
        }
      }
#pragma unroll
      for (int m = 1; m < kR; m++)
      {
        const uint32_t candidate = __byte_perm(next_pairs[m >> 1], 0u, (m & 1) ? 0x3232 : 0x1010);
#pragma unroll
        for (int p = 0; p < kPairs; p++)
        {
          if (2 * p < m)
          {
            const int d_low = kR + m - 2 * p;
            const uint32_t offsets = static_cast<uint32_t>(d_low * d_low)
                | (static_cast<uint32_t>((d_low - 1) * (d_low - 1)) << 16);
            best_pairs[p] = __viaddmin_u16x2(candidate, offsets, best_pairs[p]);
          }
        }
      }
      worst = worst_of_lane();
      // Rows from memory, four per side and round: row base - R - 1 - t below (distance
      // j + R + 1 + t from row j) and row base + 2 R + t above (distance 2 R + t - j), clamped to
      // the line (a clamped row was already seen at a smaller distance). Heights relative to the
      // chunk's class: an opposite-class row counts 0. Convergent: every lane evaluates every
      // candidate, which is harmless for certified rows (a candidate is at least (R + 1)^2).
      // Past the step allowance or the depth the 16-bit arithmetic covers the tile is given up.
      const uint32_t class_bit = chunk_filled << 31;
      const uint32_t allowance =
          (step_rate * (static_cast<uint32_t>(base - first_row) + credit_rows)) >> 7;
      int t = 0;
      while (true)
      {
        const int reach = kR + 1 + t;
        const bool open = (base - kR - 1 - t >= 0) || (base + 2 * kR + t <= last_row);
        const bool need = static_cast<uint32_t>(reach * reach) < worst && open;
        if (!__any_sync(0xffffffffu, need))
        {
          break;
        }
        if (steps > allowance || reach + 3 > deepest_search)
        {
          over_budget = true;
          break;
        }
        steps += 4u;
        constexpr int kUnroll = 4;
        uint32_t words_below[kUnroll], words_above[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++)
        {
          const int below = __viaddmax_s32(base - kR - 1 - u, -t, 0);
          const int above = __viaddmin_s32(base + 2 * kR + u, t, last_row);
          words_below[u] =
              LoadRowWord(line + static_cast<uint64_t>(static_cast<uint32_t>(below)) * stride_bytes);
          words_above[u] =
              LoadRowWord(line + static_cast<uint64_t>(static_cast<uint32_t>(above)) * stride_bytes);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++)
        {
          const uint32_t tu = static_cast<uint32_t>(t + u);
          const uint32_t tu_squared = tu * tu * 0x10001u;
          const uint32_t height_below = min(
              static_cast<uint32_t>(max(static_cast<int32_t>(words_below[u] ^ class_bit), 0)),
              kSaturated);
          const uint32_t height_above = min(
              static_cast<uint32_t>(max(static_cast<int32_t>(words_above[u] ^ class_bit), 0)),
              kSaturated);
          const uint32_t from_below = height_below * 0x10001u + tu_squared;
          const uint32_t from_above = height_above * 0x10001u + tu_squared;
          const auto fold = [&](auto pair_index)
          {
            constexpr int p = decltype(pair_index)::value;
            // (tu + a)^2 = tu^2 + 2 a tu + a^2 per half
            best_pairs[p] = FoldSquaredDistances<2 * p + kR + 1, 2 * p + kR + 2>(
                best_pairs[p], tu, from_below);
            best_pairs[p] = FoldSquaredDistances<2 * kR - 2 * p, 2 * kR - 2 * p - 1>(
                best_pairs[p], tu, from_above);
          };
          StaticFor<kPairs>(fold);
        }
        t += kUnroll;
        worst = worst_of_lane();
      }
      // A result that is still saturated says nothing (values are clamped): the tile is given up.
      if (__any_sync(0xffffffffu, worst >= kSaturated))
      {
        over_budget = true;
      }
    }
    // ---------------------------------------------------------------------------------- phase C
    if (!over_budget && !probe_only)
    {
      if constexpr (kMode == kEmitPacked)
      {
        // the class bits of this chunk's rows: row j at bit R - 1 - j
        const uint32_t chunk_only = static_cast<uint32_t>(classes >> kR) & kSideMask;
        uint32_t words[kR];
#pragma unroll
        for (int j = 0; j < kR; j++)
        {
          // the class of row j moved to bit 31 with zeros in the low half (a multiply); the
          // squared distance of an odd row comes down by a multiply too, and ONE logic
          // operation assembles the word (values stay below 0x8000: bit 15 is clear)
          const uint32_t filled = chunk_only * kPowersOfTwo[32 - kR + j];
          const uint32_t pair = best_pairs[j >> 1];
          words[j] = (j & 1) ? ((filled & 0x80000000u) | ShiftRightByMultiply<16>(pair))
                             : ((filled | pair) & 0x8000ffffu);
        }
        emit_packed(base, write_base, words, kEdge ? min(kR, last_row + 1 - base) : kR);
      }
      else
      {
        if (!kEdge && !searched)  // warp-uniform
        {
          emit_finalized(base, write_base, best_pairs, edge, std::true_type{});
        }
        else
        {
          emit_finalized(base, write_base, best_pairs, edge, std::false_type{});
        }
      }
    }
    // the next chunk becomes the current one
#pragma unroll
    for (int i = 0; i < kPairs; i++)
    {
      previous_pairs[i] = current_pairs[i];
      current_pairs[i] = next_pairs[i];
    }
    // (the class word moves on by itself: absorbing the next chunk shifts R bits in)
  };

  {
    using Edge = std::true_type;
    using Interior = std::false_type;
    // previous and current chunk of the segment's first row (a segment that does not start the
    // line reads its neighbour's rows), and the raw rows of the chunk after
#pragma unroll
    for (int i = 0; i < kR; i += 2)
    {
      const int row = first_row - kR + i;
      const uint32_t low_word = load_clamped(row);
      const uint32_t high_word = load_clamped(row + 1);
      const uint32_t low = (row < 0) ? kSaturated : clamped_value(low_word);
      const uint32_t high = (row + 1 < 0) ? kSaturated : clamped_value(high_word);
      previous_pairs[i >> 1] = __byte_perm(low, high, 0x5410);
      classes = (classes << 2) | ((low_word >> 31) << 1) | (high_word >> 31);
    }
#pragma unroll
    for (int i = 0; i < kR; i += 2)
    {
      const int row = first_row + i;
      const uint32_t low_word = load_clamped(row);
      const uint32_t high_word = load_clamped(row + 1);
      const uint32_t low = (row > last_row) ? kSaturated : clamped_value(low_word);
      const uint32_t high = (row + 1 > last_row) ? kSaturated : clamped_value(high_word);
      current_pairs[i >> 1] = __byte_perm(low, high, 0x5410);
      classes = (classes << 2) | ((low_word >> 31) << 1) | (high_word >> 31);
    }
    if constexpr (kStage != 0)
    {
      // the next chunk into buffer 0, the one after it into buffer 1
      fill_stage(0, first_row + kR, std::true_type{});
      fill_stage(1, first_row + 2 * kR, std::true_type{});
    }
    else
    {
#pragma unroll
      for (int i = 0; i < kR; i++)
      {
        raw[i] = load_clamped(first_row + kR + i);
      }
    }
    int base = first_row;
    // chunks whose own rows and whose prefetched chunk (rows base + 2 R .., kStage: + 3 R ..) lie
    // inside the line
    constexpr int kInteriorSpan = kStage ? 4 : 3;
#pragma unroll 1
    for (; base + kInteriorSpan * kR <= length && base < end_row && !over_budget; base += kR)
    {
      compute_chunk(base, Interior{});
    }
#pragma unroll 1
    for (; base < end_row && !over_budget; base += kR)
    {
      compute_chunk(base, Edge{});
    }
    // nothing of this warp may still be in flight when it exits
    if constexpr (kStage == 1)
    {
      WaitForAsyncCopies<0>();
    }
    else if constexpr (kStage == 2)
    {
#pragma unroll
      for (int b = 0; b < 3; b++)
      {
        if ((in_flight >> b) & 1u)  // warp-uniform
        {
          wait_for_stage(b);
        }
      }
    }
  }

  if (block_select == kSelectPilot)
  {
    if (lane == 0 && steps != 0u)
    {
      atomicAdd(redo + kRedoPilotSteps, steps);
    }
    if (lane == 0 && over_budget)
    {
      atomicAdd(redo + kRedoCount, 1u);
    }
    return;
  }
  if (over_budget)
  {
    // The stack kernel redoes this tile (it overwrites whatever was stored here). Several
    // segments of one tile may give up: the flag words after the list keep the entry unique.
    if (lane == 0)
    {
      const uint32_t num_tiles = tiles_per_outer * static_cast<uint32_t>(family.num_outer);
      if (atomicExch(redo + kRedoList + num_tiles + tile_index, 1u) == 0u)
      {
        const uint32_t at = atomicAdd(redo + kRedoCount, 1u);
        redo[kRedoList + at] = tile_index;
      }
    }
    return;
  }

  if constexpr (kMode != kEmitPacked)
  {
    if (min_max_keys == nullptr)
    {
      return;
    }
    using Key = typename OutputOf<kMode>::Key;
    // the lane's extrema as values: +magnitude of its deepest free voxel, -magnitude of its
    // deepest filled voxel (neutral elements where the lane emitted none of a class)
    const Out* const table = static_cast<const Out*>(finalize.magnitude_table);
    Out lane_max;
    Out lane_min;
    if constexpr (sizeof(Out) == 4)
    {
      lane_max = largest_signed >= 0 ? __int_as_float(largest_signed) : -PositiveInfinity<Out>();
      lane_min = (largest_unsigned >> 31) != 0u ? __uint_as_float(largest_unsigned)
                                                : PositiveInfinity<Out>();
    }
    else
    {
      const uint32_t deepest_free = lane_extrema & 0xffffu;
      const uint32_t deepest_filled = lane_extrema >> 16;
      lane_max = deepest_free != 0u ? __ldg(table + deepest_free) : -PositiveInfinity<Out>();
      lane_min = deepest_filled != 0u ? -__ldg(table + deepest_filled) : PositiveInfinity<Out>();
    }
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
