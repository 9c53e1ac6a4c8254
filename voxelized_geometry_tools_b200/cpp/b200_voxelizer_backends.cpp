// Replacement translation unit for the reference's backend factory
// (src/voxelized_geometry_tools/pointcloud_voxelization.cpp:18-147): it DEFINES the functions the
// reference's own header declares (include/voxelized_geometry_tools/pointcloud_voxelization.hpp:
// 53-68), so every caller of GetAvailableBackends() / MakePointCloudVoxelizer(...) keeps
// compiling and gets the sm_100a backend where it used to get CudaPointCloudVoxelizer.
//
//   BackendOptions::CUDA     -> B200PointCloudVoxelizer (b200_pointcloud_voxelization.hpp), one
//                               AvailableBackend per usable sm_100 device, option CUDA_DEVICE
//                               as in cuda_voxelization_helpers.cu:566-590
//   BackendOptions::OPENCL   -> std::runtime_error: this build carries no OpenCL backend (the
//                               reference's dummy helpers report "not available" the same way,
//                               device_pointcloud_voxelization.hpp:34-46)
//   BackendOptions::CPU      -> the reference's own CpuPointCloudVoxelizer, unchanged
//   BEST_AVAILABLE           -> CUDA (B200), else CPU
// Build: compile this file instead of pointcloud_voxelization.cpp and link libvgt_b200.so
// (INTEGRATION.md section 3).
#include <voxelized_geometry_tools/pointcloud_voxelization.hpp>

#include <memory>
#include <stdexcept>
#include <string>

#include <voxelized_geometry_tools/cpu_pointcloud_voxelization.hpp>

#include "b200_pointcloud_voxelization.hpp"
#include "vgt_b200.h"

namespace voxelized_geometry_tools
{
VGT_NAMESPACE_BEGIN
namespace pointcloud_voxelization
{
namespace
{
using VoxelizerPtr = std::unique_ptr<PointCloudVoxelizationInterface>;

VoxelizerPtr MakeB200(const std::map<std::string, int32_t>& options,
                      const LoggingFunction& logging_fn)
{
  return VoxelizerPtr(new B200PointCloudVoxelizer(options, logging_fn));
}

VoxelizerPtr MakeCpu(const std::map<std::string, int32_t>& options,
                     const LoggingFunction& logging_fn)
{
  return VoxelizerPtr(new CpuPointCloudVoxelizer(options, logging_fn));
}

void Log(const LoggingFunction& logging_fn, const std::string& message)
{
  if (logging_fn)
  {
    logging_fn(message);
  }
}
}  // namespace

std::vector<AvailableBackend> GetAvailableBackends()
{
  std::vector<AvailableBackend> backends;
  const int devices = vgt_b200_device_count();
  for (int device = 0; device < devices; device++)
  {
    backends.emplace_back("CUDA - B200 sm_100a device [" + std::to_string(device) + "]",
                          std::map<std::string, int32_t>{{"CUDA_DEVICE", device}},
                          BackendOptions::CUDA);
  }
#if defined(_OPENMP)
  backends.emplace_back("CPU/OpenMP (parallel)",
                        std::map<std::string, int32_t>{{"CPU_PARALLELIZE", 1}},
                        BackendOptions::CPU);
#else
  backends.emplace_back("CPU/async (parallel)",
                        std::map<std::string, int32_t>{{"CPU_PARALLELIZE", 1}},
                        BackendOptions::CPU);
#endif
  backends.emplace_back("CPU (serial)", std::map<std::string, int32_t>{{"CPU_PARALLELIZE", 0}},
                        BackendOptions::CPU);
  return backends;
}

std::unique_ptr<PointCloudVoxelizationInterface> MakePointCloudVoxelizer(
    const BackendOptions backend_option, const std::map<std::string, int32_t>& device_options,
    const LoggingFunction& logging_fn)
{
  switch (backend_option)
  {
    case BackendOptions::BEST_AVAILABLE:
      return MakeBestAvailablePointCloudVoxelizer(device_options, logging_fn);
    case BackendOptions::CPU:
      return MakeCpu(device_options, logging_fn);
    case BackendOptions::CUDA:
      return MakeB200(device_options, logging_fn);
    case BackendOptions::OPENCL:
      throw std::runtime_error("OpenCL PointCloud Voxelizer is not available in this build");
  }
  throw std::invalid_argument("Invalid BackendOptions");
}

std::unique_ptr<PointCloudVoxelizationInterface> MakePointCloudVoxelizer(
    const AvailableBackend& backend, const LoggingFunction& logging_fn)
{
  return MakePointCloudVoxelizer(backend.BackendOption(), backend.DeviceOptions(), logging_fn);
}

std::unique_ptr<PointCloudVoxelizationInterface> MakeBestAvailablePointCloudVoxelizer(
    const std::map<std::string, int32_t>& device_options, const LoggingFunction& logging_fn)
{
  // Preference order of the reference (pointcloud_voxelization.cpp:95-146) minus OpenCL.
  Log(logging_fn, "Trying to construct CUDA PointCloud Voxelizer...");
  try
  {
    return MakeB200(device_options, logging_fn);
  }
  catch (const std::runtime_error&)
  {
    Log(logging_fn, "CUDA PointCloud Voxelizer is not available");
  }
  Log(logging_fn, "Trying to construct CPU PointCloud Voxelizer...");
  try
  {
    return MakeCpu(device_options, logging_fn);
  }
  catch (const std::runtime_error&)
  {
    throw std::runtime_error("No PointCloud Voxelizers available");
  }
}
}  // namespace pointcloud_voxelization
VGT_NAMESPACE_END
}  // namespace voxelized_geometry_tools
