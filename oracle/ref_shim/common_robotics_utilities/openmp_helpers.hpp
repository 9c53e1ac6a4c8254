// TEST INFRASTRUCTURE ONLY -- stand-in for common_robotics_utilities/openmp_helpers.hpp: the two
// queries the reference's tests and backend factory make.
#pragma once

#include <cstdint>

#if defined(_OPENMP)
#include <omp.h>
#endif

namespace common_robotics_utilities
{
namespace openmp_helpers
{
constexpr bool IsOmpEnabledInBuild()
{
#if defined(_OPENMP)
  return true;
#else
  return false;
#endif
}

inline int32_t GetNumOmpThreads()
{
#if defined(_OPENMP)
  return static_cast<int32_t>(omp_get_max_threads());
#else
  return 1;
#endif
}
}  // namespace openmp_helpers
}  // namespace common_robotics_utilities
