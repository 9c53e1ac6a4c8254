// Device pieces of vgt_b200_edt_transform_inplace_f64: the drop-in for
// internal::ComputeDistanceFieldTransformInPlace (signed_distance_field_generation.hpp:34-37,
// .cpp:258-391) on a caller-provided double field. The envelope kernels of the SDF path take
// arbitrary partial squared distances, so the general transform is: doubles -> sign-fused words
// of ONE class (no voxel has an opposite-class neighbour, only the parabolas of the samples
// remain), three envelope passes, words -> doubles. The words hold sample + 1: in the SDF path a
// partial distance is never 0 (the nearest opposite-class voxel is at least one voxel away) and
// the stack kernels rely on that (a packed stack entry of 0 means "empty"); the transform of
// f + 1 is the transform of f, plus 1. The pass along the contiguous axis runs on a transposed
// copy (the envelope kernels want the lines strided and the lanes contiguous).
// Included by edt_kernels.cu.
#pragma once

#include "edt_kernels.cuh"

namespace vgt_b200
{
namespace edt
{
namespace
{
// Report of the ingest: [0] = samples that are neither +inf nor a non-negative integer below
// 2^31 - 2, [1] = largest finite sample + 1.
__global__ void IngestSamplesKernel(const double* __restrict__ field, int64_t count,
                                    uint32_t* __restrict__ words, uint32_t* report)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t word = kNone;
  bool bad = false;
  if (i < count)
  {
    const double sample = field[i];
    if (sample == __longlong_as_double(0x7ff0000000000000LL))
    {
      word = kNone;  // the reference's +inf: "no site here"
    }
    else if (sample >= 0.0 && sample < 2147483646.0 && sample == floor(sample))
    {
      word = static_cast<uint32_t>(sample) + 1u;
    }
    else
    {
      bad = true;
    }
    words[i] = word;
  }
  const uint32_t finite = (word == kNone) ? 0u : word;
  const uint32_t warp_max = __reduce_max_sync(0xffffffffu, finite);
  const uint32_t warp_bad = __ballot_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0)
  {
    if (warp_max != 0u)
    {
      atomicMax(report + 1, warp_max);
    }
    if (warp_bad != 0u)
    {
      atomicAdd(report + 0, static_cast<uint32_t>(__popc(warp_bad)));
    }
  }
}

__global__ void EmitSamplesKernel(const uint32_t* __restrict__ words, int64_t count,
                                  double* __restrict__ field)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < count)
  {
    const uint32_t value = words[i] & kNone;
    field[i] = (value == kNone) ? __longlong_as_double(0x7ff0000000000000LL)
                                : static_cast<double>(value - 1u);
  }
}

// out[o][c][r] = in[o][r][c] for `planes` planes of rows x columns words (32 x 32 tiles through
// shared memory, both sides coalesced).
__global__ void TransposePlanesKernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                      int rows, int columns)
{
  __shared__ uint32_t tile[32][33];
  const int64_t plane = static_cast<int64_t>(blockIdx.z) * rows * columns;
  const int column = blockIdx.x * 32 + threadIdx.x;
  for (int k = threadIdx.y; k < 32; k += blockDim.y)
  {
    const int row = blockIdx.y * 32 + k;
    if (row < rows && column < columns)
    {
      tile[k][threadIdx.x] = in[plane + static_cast<int64_t>(row) * columns + column];
    }
  }
  __syncthreads();
  const int out_column = blockIdx.y * 32 + threadIdx.x;  // a row of the input
  for (int k = threadIdx.y; k < 32; k += blockDim.y)
  {
    const int out_row = blockIdx.x * 32 + k;  // a column of the input
    if (out_row < columns && out_column < rows)
    {
      out[plane + static_cast<int64_t>(out_row) * rows + out_column] = tile[threadIdx.x][k];
    }
  }
}
}  // namespace
}  // namespace edt
}  // namespace vgt_b200
