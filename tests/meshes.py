"""Seeded triangle meshes for the rasterizer tests (vertices float64 [n, 3], triangles int32
[m, 3])."""
import numpy as np


def random_pose(rng, translation_scale=1.0):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    m = np.eye(4)
    m[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                 [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                 [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    m[:3, 3] = rng.uniform(-1, 1, 3) * translation_scale
    return m


def pose_centred_on_origin(rng, dims, resolution, jitter=0.1, max_angle=0.15):
    """A slightly rotated map origin whose grid centre lies near the world origin (where the
    meshes are), so a mesh of radius ~0.5 overlaps the map and partly leaves it. (The reference
    turns the two corners of a triangle's WORLD-frame bounding box into grid indices and loops
    from one to the other, so under a large rotation most of its loops are empty; a small angle
    keeps them populated while every rotated-frame code path still runs.)"""
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    angle = rng.uniform(-max_angle, max_angle)
    k = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    pose = np.eye(4)
    pose[:3, :3] = np.eye(3) + np.sin(angle) * k + (1 - np.cos(angle)) * (k @ k)
    centre = np.array(dims, dtype=np.float64) * resolution / 2.0
    pose[:3, 3] = -pose[:3, :3] @ centre + rng.uniform(-jitter, jitter, 3)
    return pose


def icosphere(subdivisions=2, radius=0.5, centre=(0.1, -0.2, 0.3)):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    vertices = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t),
                (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    vertices = [np.array(v, dtype=np.float64) / np.linalg.norm(v) for v in vertices]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
             (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8),
             (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdivisions):
        cache, refined = {}, []

        def midpoint(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = vertices[a] + vertices[b]
                vertices.append(m / np.linalg.norm(m))
                cache[key] = len(vertices) - 1
            return cache[key]

        for a, b, c in faces:
            ab, bc, ca = midpoint(a, b), midpoint(b, c), midpoint(c, a)
            refined += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = refined
    return (np.array(vertices) * radius + np.array(centre),
            np.array(faces, dtype=np.int32))


def box(low=(-0.3, -0.2, -0.25), high=(0.35, 0.4, 0.15)):
    lx, ly, lz = low
    hx, hy, hz = high
    vertices = np.array([[lx, ly, lz], [hx, ly, lz], [hx, hy, lz], [lx, hy, lz],
                         [lx, ly, hz], [hx, ly, hz], [hx, hy, hz], [lx, hy, hz]])
    triangles = np.array([[0, 1, 2], [0, 2, 3], [4, 6, 5], [4, 7, 6], [0, 5, 1], [0, 4, 5],
                          [1, 6, 2], [1, 5, 6], [2, 7, 3], [2, 6, 7], [3, 4, 0], [3, 7, 4]],
                         dtype=np.int32)
    return vertices, triangles


def make(name, seed=0):
    rng = np.random.default_rng(1000 + seed)
    if name == "icosphere":
        return icosphere()
    if name == "box":
        return box()
    if name == "random_soup":
        # unrelated triangles of mixed size, many with the query projecting outside them
        centres = rng.uniform(-0.5, 0.5, (60, 1, 3))
        vertices = (centres + rng.normal(size=(60, 3, 3)) * rng.uniform(0.02, 0.3, (60, 1, 1)))
        return vertices.reshape(-1, 3), np.arange(180, dtype=np.int32).reshape(60, 3)
    if name == "slivers":
        # long thin triangles (the reference warns about them) and axis-aligned ones whose
        # vertices sit exactly on cell faces
        vertices = np.array([[0.0, 0.0, 0.0], [1.0, 0.001, 0.0], [0.5, 0.0005, 0.0002],
                             [0.25, 0.25, 0.25], [0.75, 0.25, 0.25], [0.25, 0.75, 0.25],
                             [-0.5, -0.5, 0.5], [0.5, -0.5, 0.5], [0.0, 0.9, -0.4]])
        return vertices, np.array([[0, 1, 2], [3, 4, 5], [6, 7, 8]], dtype=np.int32)
    if name == "degenerate":
        # zero-area triangles: repeated vertices and collinear points (NaN / zero-normal paths)
        vertices = np.array([[0.1, 0.1, 0.1], [0.1, 0.1, 0.1], [0.4, 0.2, 0.3],
                             [0.0, 0.0, 0.0], [0.2, 0.2, 0.2], [0.4, 0.4, 0.4],
                             [-0.2, 0.3, 0.1], [-0.2, 0.3, 0.1], [-0.2, 0.3, 0.1],
                             [0.3, -0.3, 0.0], [0.5, -0.1, 0.2], [0.2, 0.0, -0.3]])
        return vertices, np.array([[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 10, 11]], dtype=np.int32)
    raise ValueError(name)
