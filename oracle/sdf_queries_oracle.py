"""CPU restatement of the reference's SignedDistanceField queries (TEST INFRASTRUCTURE ONLY).

Pure Python over numpy scalars (float64 arithmetic, float32 where the reference subtracts two
ScalarType values), one point at a time, written to follow the reference statement by statement:

    include/voxelized_geometry_tools/signed_distance_field.hpp
      :259-273   GetCorrectedCenterDistance
      :276-308   GetAxisInterpolationIndices
      :311-357   EstimateDistanceInterpolateFromNeighbors
      :823-838   EstimateLocationDistance4d
      :903-1025  GetIndexCoarseGradient / GetGridAlignedIndexCoarseGradient
      :1051-1092 GetLocationFineGradient, :213-255 ComputeAxisFineGradient
      :1159-1203 ProjectLocationOutOfCollisionToMinimumDistance4d
      :1207-1231 ComputeLocalExtremaMap, :360-480 FollowGradientsToLocalExtremaUnsafe,
                 :482-534 GradientIsEffectiveFlat / GetNextFromGradient

PARITY STATUS. The reference has no test or golden vector for any of these members (they are
exercised only by example/*.cpp), and the trilinear blend itself is
common_robotics_utilities::math::TrilinearInterpolate, which is not in the reference tree
(unvendored, unpinned). Everything else here is in-tree arithmetic mirrored operation by
operation. So: the device kernels are checked bit for bit against THIS restatement; against the
reference the blend is "parity unpinned" with a stated tolerance of 1e-12 relative (another
evaluation order of the same trilinear formula moves the last bits only). The index / location
conventions are those of oracle/ref_shim/common_robotics_utilities/voxel_grid.hpp.
"""
from __future__ import annotations

import math

import numpy as np

NO_VALUE, VALUE, THROWS = 0, 1, 2


def _inverse_rigid(m):
    from voxelized_geometry_tools_b200.grids import inverse_rigid
    return inverse_rigid(m)


class SdfOracle:
    def __init__(self, sdf: np.ndarray, resolution: float, origin_transform=None):
        self.sdf = np.ascontiguousarray(sdf, dtype=np.float32)
        self.nx, self.ny, self.nz = self.sdf.shape
        self.res = float(resolution)
        self.inv = 1.0 / self.res
        self.x_wg = np.eye(4) if origin_transform is None else \
            np.asarray(origin_transform, dtype=np.float64).reshape(4, 4)
        self.x_gw = _inverse_rigid(self.x_wg)

    # ---- frames and indices
    def _to_grid(self, p):
        m = self.x_gw
        return [((m[r, 0] * p[0] + m[r, 1] * p[1]) + m[r, 2] * p[2]) + m[r, 3] for r in range(3)]

    def _rotate_to_world(self, v):
        m = self.x_wg
        return [((m[r, 0] * v[0] + m[r, 1] * v[1]) + m[r, 2] * v[2]) + m[r, 3] * 0.0
                for r in range(3)]

    def _locate(self, p):
        q = self._to_grid(p)
        index = []
        for value in q:
            scaled = value * self.inv
            index.append(int(math.floor(scaled)) if math.isfinite(scaled) else -(2 ** 62))
        return q, index

    def _in_bounds(self, i):
        return 0 <= i[0] < self.nx and 0 <= i[1] < self.ny and 0 <= i[2] < self.nz

    def _stored(self, x, y, z):
        return float(self.sdf[x, y, z])

    # ---- distance estimate
    def _corrected(self, x, y, z):
        nominal = self._stored(x, y, z)
        offset = self.res * 0.5
        return nominal - offset if nominal >= 0.0 else nominal + offset

    @staticmethod
    def _axis_indices(initial, size, offset):
        lower = upper = initial
        if offset >= 0.0:
            upper = initial + 1
            if upper >= size:
                upper = initial
                lower = initial - 1
                if lower < 0:
                    lower = initial
        else:
            lower = initial - 1
            if lower < 0:
                upper = initial + 1
                lower = initial
                if upper >= size:
                    upper = initial
        return lower, upper

    def estimate_distance(self, p):
        q, i = self._locate(p)
        if not self._in_bounds(i):
            return NO_VALUE, 0.0
        centre = [self.res * (float(k) + 0.5) for k in i]
        x0, x1 = self._axis_indices(i[0], self.nx, q[0] - centre[0])
        y0, y1 = self._axis_indices(i[1], self.ny, q[1] - centre[1])
        z0, z1 = self._axis_indices(i[2], self.nz, q[2] - centre[2])
        low = [self.res * (float(k) + 0.5) for k in (x0, y0, z0)]

        def ratio(query, lo):
            with np.errstate(all="ignore"):
                r = np.float64(query - lo) / np.float64((lo + self.res) - lo)
            return min(max(float(r), 0.0), 1.0) if not math.isnan(r) else float("nan")

        def interpolate(a, b, r):
            with np.errstate(all="ignore"):
                return float(np.float64(a) * np.float64(1.0 - r) + np.float64(b) * np.float64(r))

        rx, ry, rz = ratio(q[0], low[0]), ratio(q[1], low[1]), ratio(q[2], low[2])
        c = self._corrected
        mm = interpolate(c(x0, y0, z0), c(x1, y0, z0), rx)
        mp = interpolate(c(x0, y0, z1), c(x1, y0, z1), rx)
        pm = interpolate(c(x0, y1, z0), c(x1, y1, z0), rx)
        pp = interpolate(c(x0, y1, z1), c(x1, y1, z1), rx)
        m = interpolate(mm, pm, ry)
        pz = interpolate(mp, pp, ry)
        return VALUE, interpolate(m, pz, rz)

    # ---- coarse gradient
    def coarse_gradient_at_index(self, x, y, z, enable_edge_gradients):
        if not self._in_bounds((x, y, z)):
            return NO_VALUE, [0.0, 0.0, 0.0]
        s = self.sdf
        with np.errstate(all="ignore"):
            if 0 < x < self.nx - 1 and 0 < y < self.ny - 1 and 0 < z < self.nz - 1:
                scale = 1.0 / (2.0 * self.res)
                # float - float is a float; then times a double
                aligned = [float(np.float64(s[x + 1, y, z] - s[x - 1, y, z]) * scale),
                           float(np.float64(s[x, y + 1, z] - s[x, y - 1, z]) * scale),
                           float(np.float64(s[x, y, z + 1] - s[x, y, z - 1]) * scale)]
            elif enable_edge_gradients:
                aligned = [0.0, 0.0, 0.0]
                for axis, (index, size) in enumerate(((x, self.nx), (y, self.ny), (z, self.nz))):
                    low, high = max(0, index - 1), min(size - 1, index + 1)
                    increment = float(high - low) * self.res
                    if increment > 0.0:
                        hi_cell, lo_cell = [x, y, z], [x, y, z]
                        hi_cell[axis], lo_cell[axis] = high, low
                        aligned[axis] = float(
                            (np.float64(s[tuple(hi_cell)]) - np.float64(s[tuple(lo_cell)]))
                            * np.float64(1.0 / increment))
            else:
                return NO_VALUE, [0.0, 0.0, 0.0]
            return VALUE, [float(v) for v in self._rotate_numpy(aligned)]

    def _rotate_numpy(self, v):
        m = self.x_wg
        v = [np.float64(c) for c in v]
        return [((m[r, 0] * v[0] + m[r, 1] * v[1]) + m[r, 2] * v[2]) + m[r, 3] * np.float64(0.0)
                for r in range(3)]

    def coarse_gradient(self, p, enable_edge_gradients=False):
        _, i = self._locate(p)
        return self.coarse_gradient_at_index(i[0], i[1], i[2], enable_edge_gradients)

    # ---- fine gradient
    def fine_gradient(self, p, nominal_window_size):
        window = abs(nominal_window_size)
        _, i = self._locate(p)
        if not self._in_bounds(i):
            return NO_VALUE, [0.0, 0.0, 0.0]
        point = self.estimate_distance(p)
        gradient = []
        for axis in range(3):
            minus_p, plus_p = list(p), list(p)
            minus_p[axis] = p[axis] - window
            plus_p[axis] = p[axis] + window
            minus, plus = self.estimate_distance(minus_p), self.estimate_distance(plus_p)
            with np.errstate(all="ignore"):
                if point[0] and minus[0] and plus[0]:
                    g = np.float64(plus[1] - minus[1]) / np.float64(plus_p[axis] - minus_p[axis])
                elif point[0] and minus[0]:
                    g = np.float64(point[1] - minus[1]) / np.float64(p[axis] - minus_p[axis])
                elif point[0] and plus[0]:
                    g = np.float64(plus[1] - point[1]) / np.float64(plus_p[axis] - p[axis])
                else:
                    return THROWS, [0.0, 0.0, 0.0]
            gradient.append(float(g))
        return VALUE, gradient

    # ---- projection
    def project_out_of_collision(self, p, minimum_distance=0.0, stepsize_multiplier=0.1,
                                 max_steps=1_000_000):
        location = [float(c) for c in p]
        _, i = self._locate(location)
        if self._in_bounds(i):
            margin = minimum_distance + self.res * stepsize_multiplier * 1e-3
            max_step = self.res * stepsize_multiplier
            distance = self.estimate_distance(location)[1]
            steps = 0
            while distance <= minimum_distance:
                status, g = self.coarse_gradient(location, True)
                if status != VALUE:
                    return NO_VALUE, [0.0, 0.0, 0.0]
                with np.errstate(all="ignore"):
                    norm = float(np.sqrt(np.float64((g[0] * g[0] + g[1] * g[1]) + g[2] * g[2])
                                         + 0.0 * 0.0))
                if not norm > self.res * 0.25:
                    return NO_VALUE, [0.0, 0.0, 0.0]
                step = min(max_step, margin - distance)
                with np.errstate(all="ignore"):
                    location = [float(np.float64(location[k])
                                      + (np.float64(g[k]) / np.float64(norm)) * np.float64(step))
                                for k in range(3)]
                status, distance = self.estimate_distance(location)
                if status != VALUE:
                    return THROWS, [0.0, 0.0, 0.0]
                steps += 1
                if steps >= max_steps:
                    return THROWS, [0.0, 0.0, 0.0]
        return VALUE, location

    # ---- local extrema map
    def _effectively_flat(self, g):
        step = self.res * 0.06125
        return abs(g[0]) <= step and abs(g[1]) <= step and abs(g[2]) <= step

    def _next_from_gradient(self, index, g):
        if self.sdf[index] < 0.0:
            g = [c * -1.0 for c in g]
        step = self.res * 0.06125
        moved = list(index)
        for axis in range(3):
            if g[axis] > step:
                moved[axis] += 1
            elif g[axis] < -step:
                moved[axis] -= 1
        return tuple(moved)

    def _centre_in_grid_frame(self, index):
        return [self.res * (float(c) + 0.5) for c in index]

    def local_extrema_map(self):
        """The sequential, memoising loop of the reference, cell by cell in storage order."""
        unset = -math.inf
        extrema = np.full((self.nx, self.ny, self.nz, 3), unset, dtype=np.float64)
        for x in range(self.nx):
            for y in range(self.ny):
                for z in range(self.nz):
                    if not np.any(extrema[x, y, z] == unset):
                        continue
                    g = self.coarse_gradient_at_index(x, y, z, True)[1]
                    if self._effectively_flat(g):
                        extrema[x, y, z] = self._centre_in_grid_frame((x, y, z))
                        continue
                    path = {(x, y, z): 1}
                    current = (x, y, z)
                    while True:
                        current = self._next_from_gradient(current, g)
                        if path.setdefault(current, 0) != 0:
                            found = self._centre_in_grid_frame(current)
                            break
                        if not self._in_bounds(current):
                            found = [math.inf, math.inf, math.inf]
                            break
                        path[current] = 1
                        if not np.any(extrema[current] == unset):
                            found = list(extrema[current])
                            break
                        g = self.coarse_gradient_at_index(*current, True)[1]
                        if self._effectively_flat(g):
                            found = self._centre_in_grid_frame(current)
                            break
                    for cell in path:
                        if self._in_bounds(cell):
                            extrema[cell] = found
        return extrema
