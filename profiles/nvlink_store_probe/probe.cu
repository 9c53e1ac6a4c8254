// Experiment (not part of the product): how fast can SMs store into a PEER's memory over NVLink,
// as a function of the bytes a warp store instruction covers? The fused exchange of the y pass
// issues one 128-byte segment per warp store (4 bytes per lane, rows 2-8 KB apart) and levels
// off at ~560 GB/s. Patterns:
//   0: 4 B per lane, consecutive warps write rows `row_stride` bytes apart (the y pass's pattern)
//   1: 4 B per lane, fully contiguous
//   2: 16 B per lane (512 B per warp store), contiguous
//   3: 16 B per lane, 512-byte pieces `row_stride` bytes apart
#include <cstdint>
#include <cuda_runtime.h>

template <int kPattern>
__global__ void StoreKernel(uint32_t* __restrict__ out, uint64_t words, uint64_t row_stride_words)
{
  const uint64_t warp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t warps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (kPattern == 0 || kPattern == 1)
  {
    const uint64_t rows = words / 32;
    for (uint64_t r = warp; r < rows; r += warps)
    {
      // pattern 0: row r lives at (r % spread) * stride + (r / spread) * 32 words
      uint64_t at = r * 32;
      if (kPattern == 0)
      {
        const uint64_t spread = words / row_stride_words;  // rows per column block
        at = (r % spread) * row_stride_words + (r / spread) * 32;
      }
      __stcs(out + at + lane, static_cast<uint32_t>(r));
    }
  }
  else
  {
    const uint64_t rows = words / 128;
    uint4* out4 = reinterpret_cast<uint4*>(out);
    for (uint64_t r = warp; r < rows; r += warps)
    {
      uint64_t at = r * 32;  // in uint4 units
      if (kPattern == 3)
      {
        const uint64_t stride4 = row_stride_words / 4;
        const uint64_t spread = (words / 4) / stride4;
        at = (r % spread) * stride4 + (r / spread) * 32;
      }
      __stcs(out4 + at + lane, make_uint4(r, r, r, r));
    }
  }
}

extern "C" int nvlink_store_probe(int pattern, void* out, uint64_t bytes, uint64_t row_stride_bytes,
                                  int blocks, int repeats, float* ms)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const uint64_t words = bytes / 4;
  const uint64_t stride_words = row_stride_bytes / 4;
  auto launch = [&]()
  {
    switch (pattern)
    {
      case 0: StoreKernel<0><<<blocks, 256>>>(static_cast<uint32_t*>(out), words, stride_words); break;
      case 1: StoreKernel<1><<<blocks, 256>>>(static_cast<uint32_t*>(out), words, stride_words); break;
      case 2: StoreKernel<2><<<blocks, 256>>>(static_cast<uint32_t*>(out), words, stride_words); break;
      default: StoreKernel<3><<<blocks, 256>>>(static_cast<uint32_t*>(out), words, stride_words); break;
    }
  };
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < repeats; i++)
  {
    launch();
  }
  cudaEventRecord(b);
  const cudaError_t status = cudaDeviceSynchronize();
  cudaEventElapsedTime(ms, a, b);
  *ms /= repeats;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return static_cast<int>(status);
}
