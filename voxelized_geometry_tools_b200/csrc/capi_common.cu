// Error text, version and device enumeration for the C-ABI (include/vgt_b200.h).
#include "common.cuh"

#include <atomic>
#include <cstring>

namespace vgt_b200
{
namespace
{
thread_local char g_last_error[512] = "";
// (a statistic, not state: nothing in the library reads it back)
std::atomic<uint64_t> g_kernel_launches{0};
}

void NoteKernelLaunch() { g_kernel_launches.fetch_add(1, std::memory_order_relaxed); }
uint64_t KernelLaunchCount() { return g_kernel_launches.load(std::memory_order_relaxed); }

void SetLastError(const char* format, ...)
{
  va_list args;
  va_start(args, format);
  vsnprintf(g_last_error, sizeof(g_last_error), format, args);
  va_end(args);
}

int FailInvalid(const char* format, ...)
{
  va_list args;
  va_start(args, format);
  vsnprintf(g_last_error, sizeof(g_last_error), format, args);
  va_end(args);
  return VGT_B200_ERR_INVALID_ARGUMENT;
}

int FailDevice(const char* what, cudaError_t error)
{
  // Same shape as the reference's CudaCheckErrors text (cuda_voxelization_helpers.cu:26-33).
  snprintf(g_last_error, sizeof(g_last_error), "[%s] Cuda error [%s]", what,
           cudaGetErrorString(error));
  // Clear the sticky "last error" so one failed call does not poison the next.
  cudaGetLastError();
  return VGT_B200_ERR_DEVICE;
}

const char* LastErrorText() { return g_last_error; }
}  // namespace vgt_b200

extern "C"
{
const char* vgt_b200_last_error(void)
{
  return vgt_b200::LastErrorText();
}

uint64_t vgt_b200_kernel_launch_count(void)
{
  return vgt_b200::KernelLaunchCount();
}

const char* vgt_b200_version(void)
{
  return "vgt_b200 0.1.0 (sm_100a)";
}

int vgt_b200_device_count(void)
{
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  int usable = 0;
  for (int device = 0; device < count; device++)
  {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) == cudaSuccess
        && major == 10)
    {
      usable++;
    }
  }
  return usable;
}
}  // extern "C"
