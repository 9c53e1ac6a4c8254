"""Shared fixtures. GPU tests are marked ``@pytest.mark.gpu``; everything else runs on CPU."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as oracle_module
    oracle_module.build()
    return oracle_module


@pytest.fixture(scope="session")
def shared_library():
    """Builds (if stale) and returns the path of libvgt_b200.so. nvcc cross-compiles on CPU."""
    from voxelized_geometry_tools_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def sdf_goldens():
    return json.loads((GOLDEN / "sdf_generation_test.json").read_text())


def occupancy_from_golden_case(case) -> tuple[np.ndarray, float]:
    """Rebuilds the occupancy grid of one reference test case from the extracted literals."""
    import math
    resolution = case["resolution"]
    shape = tuple(int(math.ceil(s / resolution)) for s in case["size_xyz"])
    occupancy = np.full(shape, case["default_occupancy"], dtype=np.float32)
    box = case["filled_box"]
    if box:
        bounds = []
        for axis, count in zip("xyz", shape):
            low, high = box[axis]
            bounds.append((low, count if high is None else high))
        (x0, x1), (y0, y1), (z0, z1) = bounds
        occupancy[x0:x1, y0:y1, z0:z1] = 1.0
    return occupancy, resolution


def random_occupancy(rng: np.random.Generator, shape, fill: float, unknown: float = 0.05,
                     blobs: bool = False) -> np.ndarray:
    if blobs:
        from scipy import ndimage
        noise = ndimage.gaussian_filter(rng.standard_normal(shape), sigma=2.0, mode="wrap")
        threshold = np.quantile(noise, 1.0 - fill) if 0.0 < fill < 1.0 else (
            -np.inf if fill >= 1.0 else np.inf)
        occupancy = (noise > threshold).astype(np.float32)
    else:
        occupancy = (rng.random(shape) < fill).astype(np.float32)
    if unknown > 0:
        occupancy[rng.random(shape) < unknown] = 0.5
    odd = rng.random(shape)
    occupancy[odd < 0.01] = 0.25
    occupancy[odd > 0.99] = 0.75
    return occupancy
