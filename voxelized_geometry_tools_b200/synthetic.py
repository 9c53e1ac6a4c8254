"""Seeded synthetic workloads for the BASELINE.json configs.

All shape decisions use integer arithmetic (half-voxel units), so the numpy (host) and torch
(device) rasterisers produce identical grids bit for bit.
"""
from __future__ import annotations

import math

import numpy as np


class Mt19937_64:
    """std::mt19937_64 (the generator the reference's tests seed with 42,
    test/voxel_raycasting_test.cpp:88-94)."""

    _N, _M = 312, 156
    _MATRIX_A = 0xB5026F5AA96619E9
    _UPPER, _LOWER = 0xFFFFFFFF80000000, 0x7FFFFFFF
    _MASK = 0xFFFFFFFFFFFFFFFF

    def __init__(self, seed: int):
        self._state = [0] * self._N
        self._state[0] = seed & self._MASK
        for i in range(1, self._N):
            previous = self._state[i - 1]
            self._state[i] = (6364136223846793005 * (previous ^ (previous >> 62)) + i) & self._MASK
        self._index = self._N

    def _twist(self) -> None:
        s = self._state
        for i in range(self._N):
            x = (s[i] & self._UPPER) | (s[(i + 1) % self._N] & self._LOWER)
            value = x >> 1
            if x & 1:
                value ^= self._MATRIX_A
            s[i] = s[(i + self._M) % self._N] ^ value
        self._index = 0

    def next_u64(self) -> int:
        if self._index >= self._N:
            self._twist()
        x = self._state[self._index]
        self._index += 1
        x ^= (x >> 29) & 0x5555555555555555
        x ^= (x << 17) & 0x71D67FFFEDA60000
        x ^= (x << 37) & 0xFFF7EEE000000000
        x ^= x >> 43
        return x & self._MASK

    def canonical(self) -> float:
        """std::generate_canonical<double, 53>: one draw, value / 2^64, clamped below 1."""
        value = float(self.next_u64()) / 18446744073709551616.0
        return math.nextafter(1.0, 0.0) if value >= 1.0 else value

    def unit(self) -> float:
        """u = (x >> 11) * 2^-53, the mapping SURVEY.md section 8d fixes for our own scenes."""
        return (self.next_u64() >> 11) * (1.0 / 9007199254740992.0)


# ------------------------------------------------------------------------------------------------
# config 1: 128^3 box scene
# ------------------------------------------------------------------------------------------------
def box_scene(n: int = 128) -> np.ndarray:
    """Tutorial / test shaped obstacles scaled to n^3 (example/tutorial.cpp:98-104,
    test/sdf_generation_test.cpp:403-409, 551-566): corner box, centre box, floor layer."""
    occupancy = np.zeros((n, n, n), dtype=np.float32)
    h, q = n // 2, n // 4
    occupancy[:h, :h, :h] = 1.0
    occupancy[q:3 * q, q:3 * q, (5 * n) // 16:(11 * n) // 16] = 1.0
    occupancy[:, :, 0] = 1.0
    return occupancy


# ------------------------------------------------------------------------------------------------
# config 2 / 4 / 5: clustered spheres
# ------------------------------------------------------------------------------------------------
def sphere_list(dims, seed: int = 42, clusters: int = 32, per_cluster: int = 24):
    """Sphere centres (in voxel units, float64 [S, 3]) and base radii (float64 [S]).

    Cluster centres uniform in the grid; members offset by 0.06*N*(2u-1) per axis; radius
    uniform in [0.015 N, 0.045 N] with N = min(dims)."""
    rng = Mt19937_64(seed)
    nx, ny, nz = dims
    n_ref = float(min(dims))
    centres, radii = [], []
    for _ in range(clusters):
        cluster = (rng.unit() * nx, rng.unit() * ny, rng.unit() * nz)
        for _ in range(per_cluster):
            centres.append(tuple(cluster[a] + 0.06 * n_ref * (2.0 * rng.unit() - 1.0)
                                 for a in range(3)))
            radii.append(n_ref * (0.015 + 0.03 * rng.unit()))
    return np.array(centres, dtype=np.float64), np.array(radii, dtype=np.float64)


def quantise_spheres(centres: np.ndarray, radii: np.ndarray, radius_scale: float = 1.0):
    """Integer form: centre in half-voxel units, 4*r^2 rounded. A voxel (i, j, k) is inside iff
    (2i+1-cx)^2 + (2j+1-cy)^2 + (2k+1-cz)^2 <= r2."""
    centres_h = np.rint(2.0 * centres).astype(np.int64)
    r2_h = np.rint(4.0 * (radii * radius_scale) ** 2).astype(np.int64)
    return centres_h, r2_h


def rasterise_spheres_numpy(dims, centres_h: np.ndarray, r2_h: np.ndarray,
                            x_range=None) -> np.ndarray:
    """``x_range`` = (x_begin, x_end) rasterises one x-slab of the grid."""
    nx, ny, nz = dims
    xb, xe = (0, nx) if x_range is None else x_range
    filled = np.zeros((xe - xb, ny, nz), dtype=bool)
    for (cx, cy, cz), r2 in zip(centres_h, r2_h):
        reach = int(math.isqrt(int(r2))) // 2 + 1
        x0, x1 = max(xb, (cx // 2) - reach), min(xe, (cx // 2) + reach + 1)
        y0, y1 = max(0, (cy // 2) - reach), min(ny, (cy // 2) + reach + 1)
        z0, z1 = max(0, (cz // 2) - reach), min(nz, (cz // 2) + reach + 1)
        if x0 >= x1 or y0 >= y1 or z0 >= z1:
            continue
        dx = (2 * np.arange(x0, x1, dtype=np.int64) + 1 - cx) ** 2
        dy = (2 * np.arange(y0, y1, dtype=np.int64) + 1 - cy) ** 2
        dz = (2 * np.arange(z0, z1, dtype=np.int64) + 1 - cz) ** 2
        inside = (dx[:, None, None] + dy[None, :, None] + dz[None, None, :]) <= r2
        filled[x0 - xb:x1 - xb, y0:y1, z0:z1] |= inside
    return filled


def rasterise_spheres_torch(dims, centres_h: np.ndarray, r2_h: np.ndarray, device,
                            x_range=None):
    """Same predicate on a torch device; ``x_range`` = (x_begin, x_end) rasterises one x-slab."""
    import torch
    nx, ny, nz = dims
    xb, xe = (0, nx) if x_range is None else x_range
    filled = torch.zeros((xe - xb, ny, nz), dtype=torch.bool, device=device)
    for (cx, cy, cz), r2 in zip(centres_h.tolist(), r2_h.tolist()):
        reach = math.isqrt(int(r2)) // 2 + 1
        x0, x1 = max(xb, (cx // 2) - reach), min(xe, (cx // 2) + reach + 1)
        y0, y1 = max(0, (cy // 2) - reach), min(ny, (cy // 2) + reach + 1)
        z0, z1 = max(0, (cz // 2) - reach), min(nz, (cz // 2) + reach + 1)
        if x0 >= x1 or y0 >= y1 or z0 >= z1:
            continue
        dx = (2 * torch.arange(x0, x1, dtype=torch.int64, device=device) + 1 - cx) ** 2
        dy = (2 * torch.arange(y0, y1, dtype=torch.int64, device=device) + 1 - cy) ** 2
        dz = (2 * torch.arange(z0, z1, dtype=torch.int64, device=device) + 1 - cz) ** 2
        inside = (dx[:, None, None] + dy[None, :, None] + dz[None, None, :]) <= r2
        filled[x0 - xb:x1 - xb, y0:y1, z0:z1] |= inside
    return filled


def _unknown_hash_numpy(index: np.ndarray) -> np.ndarray:
    """31-bit mix of the flat voxel index; a voxel is 'unknown' iff hash % 100 == 0."""
    with np.errstate(over="ignore"):
        z = index.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return ((z >> np.uint64(33)) & np.uint64(0x7FFFFFFF)).astype(np.int64)


def _unknown_hash_torch(index):
    import torch
    mask34 = (1 << 34) - 1
    mask37 = (1 << 37) - 1
    z = index * torch.tensor(-7046029254386353131, dtype=torch.int64, device=index.device)
    z = (z ^ ((z >> 30) & mask34)) * torch.tensor(-4658895280553007687, dtype=torch.int64,
                                                  device=index.device)
    z = (z ^ ((z >> 27) & mask37)) * torch.tensor(-7723592293110705685, dtype=torch.int64,
                                                  device=index.device)
    return (z >> 33) & 0x7FFFFFFF


def fit_radius_scale(dims, centres, radii, target_fill: float = 0.10, tolerance: float = 0.005,
                     probe_dims=None) -> float:
    """Bisection on one radius multiplier until the filled fraction hits the target. The search
    runs on a reduced copy of the scene (same relative geometry) to stay cheap for big grids."""
    if probe_dims is None:
        shrink = max(1, min(dims) // 128)
        probe_dims = tuple(max(1, d // shrink) for d in dims)
    factor = np.array([p / d for p, d in zip(probe_dims, dims)])
    probe_centres = centres * factor[None, :]
    probe_radii = radii * float(factor.min())
    low, high = 0.2, 3.0
    scale = 1.0
    for _ in range(24):
        scale = 0.5 * (low + high)
        c_h, r2_h = quantise_spheres(probe_centres, probe_radii, scale)
        fill = float(rasterise_spheres_numpy(probe_dims, c_h, r2_h).mean())
        if abs(fill - target_fill) <= tolerance * 0.5:
            break
        if fill < target_fill:
            low = scale
        else:
            high = scale
    return scale


def clustered_spheres_occupancy(dims, seed: int = 42, target_fill: float = 0.10,
                                unknown_percent: bool = True, x_range=None) -> np.ndarray:
    """Config 2 (host): ~10 % filled clustered spheres, 1 % of voxels set to 0.5 (unknown);
    optionally one x-slab of the grid."""
    dims = tuple(int(d) for d in dims)
    centres, radii = sphere_list(dims, seed)
    scale = fit_radius_scale(dims, centres, radii, target_fill)
    c_h, r2_h = quantise_spheres(centres, radii, scale)
    filled = rasterise_spheres_numpy(dims, c_h, r2_h, x_range)
    occupancy = filled.astype(np.float32)
    if unknown_percent:
        first = 0 if x_range is None else x_range[0] * dims[1] * dims[2]
        flat = occupancy.reshape(-1)
        chunk = 1 << 24
        for start in range(0, flat.size, chunk):
            index = np.arange(start, min(flat.size, start + chunk), dtype=np.int64)
            flat[start:start + index.size][_unknown_hash_numpy(index + first) % 100 == 0] = 0.5
    return occupancy


def clustered_spheres_occupancy_torch(dims, device, seed: int = 42, target_fill: float = 0.10,
                                      unknown_percent: bool = True, x_range=None):
    """Config 2/4/5 (device): the same grid generated on the GPU, optionally one x-slab of it."""
    import torch
    dims = tuple(int(d) for d in dims)
    centres, radii = sphere_list(dims, seed)
    scale = fit_radius_scale(dims, centres, radii, target_fill)
    c_h, r2_h = quantise_spheres(centres, radii, scale)
    filled = rasterise_spheres_torch(dims, c_h, r2_h, device, x_range)
    occupancy = filled.to(torch.float32)
    del filled
    if unknown_percent:
        xb = 0 if x_range is None else x_range[0]
        plane = dims[1] * dims[2]
        flat = occupancy.view(-1)
        chunk = 1 << 26
        for start in range(0, flat.numel(), chunk):
            stop = min(flat.numel(), start + chunk)
            index = torch.arange(start, stop, dtype=torch.int64, device=device) + xb * plane
            flat[start:stop][_unknown_hash_torch(index) % 100 == 0] = 0.5
    return occupancy


def terrain_occupancy_torch(dims, device, seed: int = 7, x_range=None):
    """Config 5 terrain: filled iff z < h(x, y), h = 128 + sum A_k sin(2 pi (f_k x + g_k y)/4096 + phi_k)."""
    import torch
    nx, ny, nz = dims
    rng = Mt19937_64(seed)
    xb, xe = (0, nx) if x_range is None else x_range
    x = torch.arange(xb, xe, dtype=torch.float64, device=device)[:, None]
    y = torch.arange(0, ny, dtype=torch.float64, device=device)[None, :]
    height = torch.full((xe - xb, ny), 128.0 * nz / 512.0, dtype=torch.float64, device=device)
    for _ in range(6):
        amplitude = (8.0 + 24.0 * rng.unit()) * nz / 512.0
        fx = float(int(1 + 12 * rng.unit()))
        fy = float(int(1 + 12 * rng.unit()))
        phase = 2.0 * math.pi * rng.unit()
        height += amplitude * torch.sin(2.0 * math.pi * (fx * x / nx + fy * y / ny) + phase)
    height = height.clamp(1.0, nz - 2.0).floor().to(torch.int64)
    z = torch.arange(0, nz, dtype=torch.int64, device=device)[None, None, :]
    return (z < height[:, :, None]).to(torch.float32)


# ------------------------------------------------------------------------------------------------
# config 3: depth cameras looking at a sphere scene
# ------------------------------------------------------------------------------------------------
def optical_from_physical() -> np.ndarray:
    """X_CO of test/pointcloud_voxelization_test.cpp:192-194: Rz(-90 deg) * Rx(-90 deg)."""
    rz = np.array([[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    rx = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, -1.0, 0.0]])
    transform = np.eye(4)
    transform[:3, :3] = rz @ rx
    return transform


def look_at_pose(position, target=(0.0, 0.0, 0.0)) -> np.ndarray:
    """Physical camera frame: x forward (towards target), z up-ish."""
    position = np.asarray(position, dtype=np.float64)
    forward = np.asarray(target, dtype=np.float64) - position
    forward /= np.linalg.norm(forward)
    up = np.array([0.0, 0.0, 1.0])
    left = np.cross(up, forward)
    left /= np.linalg.norm(left)
    true_up = np.cross(forward, left)
    pose = np.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = forward, left, true_up, position
    return pose


def depth_camera_scene(grid_n: int = 256, voxel_size: float = 0.02, width: int = 640,
                       height: int = 480, seed: int = 43, num_cameras: int = 4,
                       max_range: float = 6.0, nan_fraction: float = 0.005):
    """Config 3. Returns dict(static_occupancy, origin_transform, clouds=[(points, X_WC, max_range)]).

    Scene: floor plane z = z_floor plus 64 spheres in world coordinates; pinhole cameras
    (fx = fy = 525, cx = 319.5, cy = 239.5 at 640x480, scaled with the image) at (+-4, 0, 1.5),
    (0, +-4, 1.5) looking at the origin; depth = first hit, else 8 m; 0.5 % of pixels NaN."""
    rng = Mt19937_64(seed)
    half = 0.5 * grid_n * voxel_size
    origin_transform = np.eye(4)
    origin_transform[:3, 3] = (-half, -half, -half)
    static = np.zeros((grid_n, grid_n, grid_n), dtype=np.float32)
    static[:, :, 0] = 1.0
    z_floor = -half + voxel_size  # top of the filled floor layer

    spheres = []
    for _ in range(8):
        cluster = np.array([(2.0 * rng.unit() - 1.0) * 0.7 * half,
                            (2.0 * rng.unit() - 1.0) * 0.7 * half,
                            z_floor + rng.unit() * 0.8 * half])
        for _ in range(8):
            centre = cluster + 0.12 * half * np.array(
                [2.0 * rng.unit() - 1.0, 2.0 * rng.unit() - 1.0, 2.0 * rng.unit() - 1.0])
            spheres.append((centre, half * (0.03 + 0.06 * rng.unit())))
    centres = np.array([s[0] for s in spheres])
    radii = np.array([s[1] for s in spheres])

    fx = 525.0 * width / 640.0
    fy = 525.0 * height / 480.0
    cx = (width - 1) / 2.0
    cy = (height - 1) / 2.0
    u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    # optical frame: z forward, x right, y down
    directions = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], axis=-1).reshape(-1, 3)
    x_co = optical_from_physical()
    positions = [(4.0, 0.0, 1.5), (-4.0, 0.0, 1.5), (0.0, 4.0, 1.5), (0.0, -4.0, 1.5)]
    scale = half / 2.56  # positions are quoted for the 5.12 m cube of config 3
    clouds = []
    for cam in range(num_cameras):
        position = np.array(positions[cam % 4]) * scale
        x_wc = look_at_pose(position) @ x_co
        rotation = x_wc[:3, :3]
        world_dirs = directions @ rotation.T
        norms = np.linalg.norm(world_dirs, axis=1)
        unit_dirs = world_dirs / norms[:, None]
        # t along the unit ray; depth along optical z = t / norm
        t_hit = np.full(unit_dirs.shape[0], np.inf)
        # floor
        dz = unit_dirs[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            t_floor = (z_floor - position[2]) / dz
        valid = (dz < 0) & (t_floor > 0)
        t_hit = np.where(valid, np.minimum(t_hit, t_floor), t_hit)
        for centre, radius in zip(centres, radii):
            oc = position - centre
            b = unit_dirs @ oc
            c = oc @ oc - radius * radius
            disc = b * b - c
            root = np.sqrt(np.where(disc > 0, disc, 0.0))
            t0 = -b - root
            hit = (disc > 0) & (t0 > 0)
            t_hit = np.where(hit, np.minimum(t_hit, t0), t_hit)
        depth = np.where(np.isfinite(t_hit), t_hit / norms, 8.0 * scale)
        points = directions * depth[:, None]
        pixel = np.arange(points.shape[0], dtype=np.int64) + cam * points.shape[0]
        period = max(1, int(round(1.0 / nan_fraction))) if nan_fraction > 0 else 0
        selector = (_unknown_hash_numpy(pixel) % period == 0) if period else np.zeros(
            points.shape[0], dtype=bool)
        points[selector] = np.nan
        clouds.append((np.ascontiguousarray(points), x_wc, max_range * scale))
    return {"static_occupancy": static, "origin_transform": origin_transform,
            "voxel_size": voxel_size, "clouds": clouds}
