// =============================================================================
// TEST INFRASTRUCTURE ONLY -- CPU oracle for the occupancy -> SDF path.
//
// This file is a from-scratch CPU restatement of the reference's algorithm
// (calderpg/voxelized_geometry_tools). It exists so that tests/, bench.py's
// cpu_baseline / --impl reference legs and __graft_entry__.smoke() can check
// the CUDA path. Nothing in voxelized_geometry_tools_b200/ may link, import or
// call it: the product path has no CPU fallback.
//
// Parity status: PINNED for the SDF path against every known-answer value in
// the reference's test/sdf_generation_test.cpp (tests/golden/sdf_generation_test.json,
// extracted by tests/golden/extract_reference_goldens.py) and, when /root/reference
// is present, against the reference's own signed_distance_field_generation.cpp
// compiled unmodified over shim headers (oracle/_ref, see oracle/Makefile).
//
// What it follows (paths relative to the reference checkout):
//   include/voxelized_geometry_tools/signed_distance_field_generation.hpp:39-113
//       two double fields (+inf), mark 0.0, two transforms, sqrt*res combine
//   include/voxelized_geometry_tools/signed_distance_field_generation.hpp:115-285
//       virtual-border mode: two enlarged extractions, three-way merge
//   src/voxelized_geometry_tools/signed_distance_field_generation.cpp:85-122
//       O(n^2) line transform for n <= 8
//   src/voxelized_geometry_tools/signed_distance_field_generation.cpp:124-226
//       Felzenszwalb-Huttenlocher lower envelope with inf-safe subtraction
//   src/voxelized_geometry_tools/signed_distance_field_generation.cpp:258-391
//       pass order X, Y, Z; an axis with one cell is skipped; static split of
//       lines over threads
//   include/voxelized_geometry_tools/occupancy_map.hpp:181-205
//       filled predicate: occ > 0.5, or occ == 0.5 when unknown_is_filled
//   include/voxelized_geometry_tools/signed_distance_field.hpp:765-787
//       Lock(): min/max over all cells
//
// Storage convention (common_robotics_utilities VoxelGrid, mirrored in-tree at
// src/voxelized_geometry_tools/cuda_voxelization_helpers.cu:281-282):
//   linear index = x * (ny * nz) + y * nz + z   (x slowest, z contiguous)
//
// Build: g++ -O3 -march=native -fopenmp -ffp-contract=off (see oracle/Makefile).
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace
{
constexpr double kPosInf = std::numeric_limits<double>::infinity();

// The reference switches from the quadratic scan to the envelope at n > 8
// (signed_distance_field_generation.cpp:236).
constexpr int64_t kQuadraticLimit = 8;

struct GridDims
{
  int64_t nx;
  int64_t ny;
  int64_t nz;
  int64_t Total() const { return nx * ny * nz; }
  int64_t Flat(int64_t x, int64_t y, int64_t z) const
  {
    return (x * ny + y) * nz + z;
  }
};

int ResolveThreads(int requested)
{
#ifdef _OPENMP
  if (requested <= 0)
  {
    return omp_get_max_threads();
  }
  return requested;
#else
  (void)requested;
  return 1;
#endif
}

// One line of the grid, addressed in place with a stride the way the
// reference's X/Y/Z indexers address the grid (sdfgen.cpp:37-83).
struct StridedLine
{
  double* first;
  int64_t step;
  double At(int64_t i) const { return first[i * step]; }
  void Put(int64_t i, double value) const { first[i * step] = value; }
};

struct LineScratch
{
  std::vector<double> breakpoints;  // "z" in the paper, n + 1 entries
  std::vector<int64_t> sites;       // "v" in the paper, n entries
  std::vector<double> result;       // "d" in the paper, n entries
  explicit LineScratch(int64_t n)
      : breakpoints(static_cast<size_t>(n + 1)),
        sites(static_cast<size_t>(n)),
        result(static_cast<size_t>(n)) {}
};

inline double SquareOf(int64_t value)
{
  return static_cast<double>(value * value);
}

// sdfgen.cpp:85-122 -- every output scans every input.
void QuadraticLineTransform(const StridedLine& line, int64_t n, LineScratch& s)
{
  std::fill(s.result.begin(), s.result.begin() + n, kPosInf);
  for (int64_t target = 0; target < n; target++)
  {
    for (int64_t source = 0; source < n; source++)
    {
      const double candidate = SquareOf(target - source) + line.At(source);
      if (candidate < s.result[target])
      {
        s.result[target] = candidate;
      }
    }
  }
  for (int64_t i = 0; i < n; i++)
  {
    line.Put(i, s.result[i]);
  }
}

// sdfgen.cpp:153-172 -- a - b that never produces NaN.
inline double DifferenceWithoutNaN(double a, double b)
{
  const bool a_inf = (a == kPosInf);
  const bool b_inf = (b == kPosInf);
  if (a_inf && b_inf) return 0.0;
  if (a_inf) return kPosInf;
  if (b_inf) return -kPosInf;
  return a - b;
}

// sdfgen.cpp:124-226 -- lower envelope of parabolas, then a sweep.
void EnvelopeLineTransform(const StridedLine& line, int64_t n, LineScratch& s)
{
  std::fill(s.breakpoints.begin(), s.breakpoints.end(), 0.0);
  std::fill(s.sites.begin(), s.sites.end(), int64_t{0});
  std::fill(s.result.begin(), s.result.end(), 0.0);
  s.breakpoints[0] = -kPosInf;
  s.breakpoints[1] = kPosInf;

  // Crossing abscissa of the parabola rooted at q with the one at sites[k]
  // (sdfgen.cpp:175-184).
  const auto crossing = [&](int64_t q, int64_t k)
  {
    const int64_t vk = s.sites[k];
    const double numerator = DifferenceWithoutNaN(
        line.At(q) + SquareOf(q), line.At(vk) + SquareOf(vk));
    const double denominator = static_cast<double>((2 * q) - (2 * vk));
    return numerator / denominator;
  };

  int64_t top = 0;
  for (int64_t q = 1; q < n; q++)
  {
    double x = crossing(q, top);
    while (top > 0 && x <= s.breakpoints[top])
    {
      top--;
      x = crossing(q, top);
    }
    top++;
    s.sites[top] = q;
    s.breakpoints[top] = x;
    s.breakpoints[top + 1] = kPosInf;
  }

  int64_t cursor = 0;
  for (int64_t q = 0; q < n; q++)
  {
    while (s.breakpoints[cursor + 1] < static_cast<double>(q))
    {
      cursor++;
    }
    const int64_t vk = s.sites[cursor];
    s.result[q] = SquareOf(q - vk) + line.At(vk);
  }
  for (int64_t q = 0; q < n; q++)
  {
    line.Put(q, s.result[q]);
  }
}

inline void LineTransform(const StridedLine& line, int64_t n, LineScratch& s)
{
  if (n > kQuadraticLimit)
  {
    EnvelopeLineTransform(line, n, s);
  }
  else
  {
    QuadraticLineTransform(line, n, s);
  }
}

// Runs `count` independent lines split statically over threads
// (StaticParallelForRangeLoop in the reference, sdfgen.cpp:308-311).
template <typename LineOf>
void ForEachLine(int64_t count, int64_t length, int threads, LineOf line_of)
{
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
  {
    LineScratch scratch(length);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < count; i++)
    {
      LineTransform(line_of(i), length, scratch);
    }
  }
#else
  (void)threads;
  LineScratch scratch(length);
  for (int64_t i = 0; i < count; i++)
  {
    LineTransform(line_of(i), length, scratch);
  }
#endif
}

// sdfgen.cpp:258-391 -- X lines, then Y lines, then Z lines, in place.
void SquaredDistanceTransformInPlace(
    double* field, const GridDims& g, int threads)
{
  const int64_t plane = g.ny * g.nz;
  if (g.nx > 1)
  {
    ForEachLine(g.ny * g.nz, g.nx, threads, [&](int64_t i)
    {
      const int64_t y = i / g.nz;
      const int64_t z = i % g.nz;
      return StridedLine{field + g.Flat(0, y, z), plane};
    });
  }
  if (g.ny > 1)
  {
    ForEachLine(g.nx * g.nz, g.ny, threads, [&](int64_t i)
    {
      const int64_t x = i / g.nz;
      const int64_t z = i % g.nz;
      return StridedLine{field + g.Flat(x, 0, z), g.nz};
    });
  }
  if (g.nz > 1)
  {
    ForEachLine(g.nx * g.ny, g.nz, threads, [&](int64_t i)
    {
      const int64_t x = i / g.ny;
      const int64_t y = i % g.ny;
      return StridedLine{field + g.Flat(x, y, 0), int64_t{1}};
    });
  }
}

using FilledPredicate = std::function<bool(int64_t, int64_t, int64_t)>;

struct SquaredFields
{
  std::vector<double> to_filled;
  std::vector<double> to_free;
};

// sdfgen.hpp:47-80 -- serial marking loop, then one transform per field.
SquaredFields ComputeSquaredFields(
    const GridDims& g, const FilledPredicate& is_filled, int threads)
{
  SquaredFields fields;
  fields.to_filled.assign(static_cast<size_t>(g.Total()), kPosInf);
  fields.to_free.assign(static_cast<size_t>(g.Total()), kPosInf);
  for (int64_t x = 0; x < g.nx; x++)
  {
    for (int64_t y = 0; y < g.ny; y++)
    {
      for (int64_t z = 0; z < g.nz; z++)
      {
        const size_t at = static_cast<size_t>(g.Flat(x, y, z));
        if (is_filled(x, y, z))
        {
          fields.to_filled[at] = 0.0;
        }
        else
        {
          fields.to_free[at] = 0.0;
        }
      }
    }
  }
  SquaredDistanceTransformInPlace(fields.to_filled.data(), g, threads);
  SquaredDistanceTransformInPlace(fields.to_free.data(), g, threads);
  return fields;
}

// sdfgen.hpp:39-113 -- the plain (no border) extraction.
template <typename Scalar>
void ExtractPlain(
    const GridDims& g, const FilledPredicate& is_filled, double resolution,
    int threads, Scalar* sdf)
{
  const SquaredFields fields = ComputeSquaredFields(g, is_filled, threads);
  for (int64_t x = 0; x < g.nx; x++)
  {
    for (int64_t y = 0; y < g.ny; y++)
    {
      for (int64_t z = 0; z < g.nz; z++)
      {
        const size_t at = static_cast<size_t>(g.Flat(x, y, z));
        const double outside = std::sqrt(fields.to_filled[at]) * resolution;
        const double inside = std::sqrt(fields.to_free[at]) * resolution;
        sdf[at] = static_cast<Scalar>(outside - inside);
      }
    }
  }
}

// sdfgen.hpp:134-284 -- the literal enlarged-grid algorithm.
template <typename Scalar>
void ExtractWithVirtualBorder(
    const GridDims& g, const FilledPredicate& is_filled, double resolution,
    int threads, Scalar* sdf)
{
  const int64_t pad_x = (g.nx > 1) ? 1 : 0;
  const int64_t pad_y = (g.ny > 1) ? 1 : 0;
  const int64_t pad_z = (g.nz > 1) ? 1 : 0;
  const GridDims big{g.nx + 2 * pad_x, g.ny + 2 * pad_y, g.nz + 2 * pad_z};

  const auto on_shell = [&](int64_t x, int64_t y, int64_t z)
  {
    return (pad_x && (x == 0 || x == big.nx - 1))
        || (pad_y && (y == 0 || y == big.ny - 1))
        || (pad_z && (z == 0 || z == big.nz - 1));
  };
  const auto with_shell_value = [&](bool shell_value)
  {
    return FilledPredicate([&, shell_value](int64_t x, int64_t y, int64_t z)
    {
      if (on_shell(x, y, z))
      {
        return shell_value;
      }
      return is_filled(x - pad_x, y - pad_y, z - pad_z);
    });
  };

  std::vector<Scalar> shell_filled(static_cast<size_t>(big.Total()));
  std::vector<Scalar> shell_empty(static_cast<size_t>(big.Total()));
  ExtractPlain<Scalar>(
      big, with_shell_value(true), resolution, threads, shell_filled.data());
  ExtractPlain<Scalar>(
      big, with_shell_value(false), resolution, threads, shell_empty.data());

  for (int64_t x = 0; x < g.nx; x++)
  {
    for (int64_t y = 0; y < g.ny; y++)
    {
      for (int64_t z = 0; z < g.nz; z++)
      {
        const size_t from =
            static_cast<size_t>(big.Flat(x + pad_x, y + pad_y, z + pad_z));
        const Scalar free_side = shell_filled[from];
        const Scalar filled_side = shell_empty[from];
        Scalar merged;
        if (free_side >= 0.0)
        {
          merged = free_side;
        }
        else if (filled_side <= -0.0)
        {
          merged = filled_side;
        }
        else
        {
          merged = static_cast<Scalar>(0.0);
        }
        sdf[static_cast<size_t>(g.Flat(x, y, z))] = merged;
      }
    }
  }
}

// signed_distance_field.hpp:765-787 -- serial min/max of the raw data.
template <typename Scalar>
void MinMaxOf(const Scalar* data, int64_t count, Scalar* min_max)
{
  if (count <= 0)
  {
    return;
  }
  const auto extrema = std::minmax_element(data, data + count);
  min_max[0] = *extrema.first;
  min_max[1] = *extrema.second;
}

template <typename Scalar>
int ExtractSdf(
    const GridDims& g, const FilledPredicate& is_filled, double resolution,
    int add_virtual_border, int threads, Scalar* sdf, Scalar* min_max)
{
  if (g.nx < 1 || g.ny < 1 || g.nz < 1 || sdf == nullptr)
  {
    return 1;
  }
  const int resolved = ResolveThreads(threads);
  if (add_virtual_border)
  {
    ExtractWithVirtualBorder<Scalar>(g, is_filled, resolution, resolved, sdf);
  }
  else
  {
    ExtractPlain<Scalar>(g, is_filled, resolution, resolved, sdf);
  }
  if (min_max != nullptr)
  {
    MinMaxOf<Scalar>(sdf, g.Total(), min_max);
  }
  return 0;
}

// occupancy_map.hpp:181-205. The float is promoted to double against the
// literal 0.5, which is exact.
FilledPredicate OccupancyPredicate(
    const float* occupancy, const GridDims& g, int unknown_is_filled)
{
  return FilledPredicate([=](int64_t x, int64_t y, int64_t z)
  {
    const double occ = static_cast<double>(occupancy[g.Flat(x, y, z)]);
    if (occ > 0.5)
    {
      return true;
    }
    if (unknown_is_filled && occ == 0.5)
    {
      return true;
    }
    return false;
  });
}

FilledPredicate MaskPredicate(const uint8_t* mask, const GridDims& g)
{
  return FilledPredicate([=](int64_t x, int64_t y, int64_t z)
  {
    return mask[g.Flat(x, y, z)] != 0;
  });
}
}  // namespace

extern "C"
{
int vgt_oracle_max_threads(void)
{
  return ResolveThreads(0);
}

// Squared distance fields in voxel units, +inf where the set is empty.
int vgt_oracle_edt_sq_f64(
    const float* occupancy, int64_t nx, int64_t ny, int64_t nz,
    int unknown_is_filled, int threads, double* dist_to_filled_sq,
    double* dist_to_free_sq)
{
  const GridDims g{nx, ny, nz};
  if (nx < 1 || ny < 1 || nz < 1)
  {
    return 1;
  }
  const SquaredFields fields = ComputeSquaredFields(
      g, OccupancyPredicate(occupancy, g, unknown_is_filled),
      ResolveThreads(threads));
  std::memcpy(dist_to_filled_sq, fields.to_filled.data(),
              sizeof(double) * fields.to_filled.size());
  std::memcpy(dist_to_free_sq, fields.to_free.data(),
              sizeof(double) * fields.to_free.size());
  return 0;
}

// In-place transform of a caller-provided field (0 / +inf or any non-negative
// samples), mirroring ComputeDistanceFieldTransformInPlace (sdfgen.hpp:34-37).
int vgt_oracle_transform_inplace_f64(
    double* field, int64_t nx, int64_t ny, int64_t nz, int threads)
{
  if (nx < 1 || ny < 1 || nz < 1 || field == nullptr)
  {
    return 1;
  }
  SquaredDistanceTransformInPlace(
      field, GridDims{nx, ny, nz}, ResolveThreads(threads));
  return 0;
}

int vgt_oracle_sdf_f32(
    const float* occupancy, int64_t nx, int64_t ny, int64_t nz,
    double resolution, int unknown_is_filled, int add_virtual_border,
    int threads, float* sdf, float* min_max)
{
  const GridDims g{nx, ny, nz};
  return ExtractSdf<float>(
      g, OccupancyPredicate(occupancy, g, unknown_is_filled), resolution,
      add_virtual_border, threads, sdf, min_max);
}

int vgt_oracle_sdf_f64(
    const float* occupancy, int64_t nx, int64_t ny, int64_t nz,
    double resolution, int unknown_is_filled, int add_virtual_border,
    int threads, double* sdf, double* min_max)
{
  const GridDims g{nx, ny, nz};
  return ExtractSdf<double>(
      g, OccupancyPredicate(occupancy, g, unknown_is_filled), resolution,
      add_virtual_border, threads, sdf, min_max);
}

int vgt_oracle_sdf_from_mask_f32(
    const uint8_t* filled_mask, int64_t nx, int64_t ny, int64_t nz,
    double resolution, int add_virtual_border, int threads, float* sdf,
    float* min_max)
{
  const GridDims g{nx, ny, nz};
  return ExtractSdf<float>(
      g, MaskPredicate(filled_mask, g), resolution, add_virtual_border,
      threads, sdf, min_max);
}
}  // extern "C"
