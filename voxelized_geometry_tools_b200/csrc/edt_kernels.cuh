// Exact signed squared Euclidean distance transform for sm_100a, as three line passes.
//
// Representation between passes ("sign-fused", one int32 per voxel for BOTH reference fields):
//   bit 31      class of the voxel (1 = filled, 0 = free)
//   bits 0..30  partial squared distance (voxel units) to the nearest voxel of the OPPOSITE class
//               over the axes processed so far; kNone (0x7fffffff) = none found yet.
// The reference keeps two VoxelGrid<double> fields (sdfgen.hpp:47-55); at every voxel exactly one
// of them is 0, so the non-zero one plus the class bit carries the same information.
//
// Why one sweep per line serves both fields: a voxel of the opposite class is a zero-height site,
// and a zero-height site hides every site behind it. So a line splits into maximal same-class
// runs; inside a run [a, b] the answer for q is
//     min( lower envelope of the run's own parabolas at q, (q-(a-1))^2, ((b+1)-q)^2 )
// with the two boundary terms present only when a-1 / b+1 are inside the line.
//
// The 1-D transform is exact (squared distances are integers; the pop test is a cross-multiplied
// integer comparison), so the result equals the reference's F-H/brute-force output
// (sdfgen.cpp:85-226) independent of pass order.
#pragma once

#include <cstdint>

namespace vgt_b200
{
namespace edt
{
constexpr uint32_t kNone = 0x7fffffffu;      // "+inf" in bits 0..30
constexpr uint32_t kClassBit = 0x80000000u;  // filled
constexpr int kWarp = 32;

// Geometry of one family of parallel lines. Column c in [0, inner_count) of outer block o starts
// at element o * outer_stride + c; consecutive elements of a line are line_stride apart.
struct LineFamily
{
  int64_t num_outer;
  int64_t outer_stride;
  int64_t inner_count;
  int64_t line_stride;
  int32_t length;
  // Optional "send layout" for the output of the slab-local y pass (multi-GPU): the line axis is
  // cut into out_parts near-equal parts (the first out_extra parts hold out_base + 1 rows, the
  // rest out_base) and part h of every line is stored as the block [num_outer][rows_h][inner],
  // blocks back to back. That is exactly what the all-to-all sends to rank h, so no packing copy
  // is needed. out_parts == 0: the output has the input's layout.
  int32_t out_parts;
  int32_t out_base;
  int32_t out_extra;
  // Fused exchange (multi-GPU): when scatter_base[0] != nullptr, part h is not written into the
  // local send buffer but straight into rank h's receive buffer [nx_total][rows_h][inner] through
  // its peer-mapped pointer scatter_base[h] (NVLink stores), at rows scatter_row_offset + outer.
  // The y pass then IS the all-to-all: no collective call, no intermediate copy.
  int64_t scatter_row_offset;
  uint32_t* scatter_base[8];
  // Fused exchange: this rank's index. The window kernel starts its sweep over the line segments
  // at the part that belongs to rank scatter_rank + 1, so that at any moment the ranks store
  // into DIFFERENT peers (all ranks sweeping the parts in the same order would aim every NVLink
  // store of the box at one receiver at a time).
  int32_t scatter_rank;
  uint32_t first_segment;     // derived by the launcher from scatter_rank
  // Derived values, filled by the launcher (FillDerived) so that the kernels read them straight
  // from the constant bank as instruction operands instead of re-deriving them in the hot loops.
  uint32_t stride_bytes;      // line_stride * 4 (the packed intermediate)
  uint32_t out_stride_bytes;  // line_stride * sizeof(output element)
  uint32_t last_row;          // length - 1
  uint32_t num_words;         // ceil(length / 32)
};

// What the last pass needs to turn squared voxel distances into the SDF.
struct FinalizeParams
{
  double resolution;
  int32_t add_virtual_border;
  // Full-grid extents and the position of this (slab) family inside the full grid, used only by
  // the virtual border: voxel (line index q, column c) is at
  //   x = q, y = y_offset + c / nz, z = c % nz.
  int32_t nx_total;
  int32_t ny_total;
  int32_t nz_total;
  int32_t y_offset;
  int32_t nz;  // columns per y row of this family
  // Optional table of the finished magnitudes, magnitude_table[s] = (Out)(sqrt((double)s) *
  // resolution) for s < magnitude_table_size, built per call by BuildMagnitudeTableKernel with
  // the very expression of the direct path: one (mostly L1-resident) load replaces the fp64
  // square root for every squared distance below the table size. nullptr / 0: no table.
  const void* magnitude_table;
  uint32_t magnitude_table_size;
};
}  // namespace edt
}  // namespace vgt_b200
