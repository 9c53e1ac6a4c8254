"""CPU: the REFERENCE'S OWN test files (test/sdf_generation_test.cpp, mesh_rasterization_test.cpp,
voxel_raycasting_test.cpp, pointcloud_voxelization_test.cpp), compiled unmodified against the reference's own sources over the
stand-in third-party layer of oracle/ref_shim and a minimal stand-in for googletest
(`make -C oracle ref_tests`, where /root/reference exists; the binaries travel with the repo).

What it says: the layer under the reference's code that this repo had to restate (Eigen,
common_robotics_utilities: grid indexing, sizes, the parallel-for helpers, the SDF container's
base) behaves, under the reference's own assertions, the way the reference expects - 16 + 4 + 1
+ 2 tests, every SDF known answer of the reference included; the voxelization test goes through
the reference's own backend factory (its dummy CUDA / OpenCL helpers leave the CPU backends). The oracle library is built from the
same sources over the same stand-ins."""
import re
import subprocess
from pathlib import Path

import pytest

BINARIES = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "tests"
EXPECTED = {"sdf_generation_test": 16, "mesh_rasterization_test": 4, "voxel_raycasting_test": 1,
            "pointcloud_voxelization_test": 2}


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_the_references_own_tests_pass_on_the_oracle_build(name):
    binary = BINARIES / name
    if not binary.exists():
        pytest.skip(f"{name} not built (make -C oracle ref_tests needs /root/reference)")
    result = subprocess.run([str(binary)], capture_output=True, text=True, timeout=900)
    summary = re.search(r"\[ DONE \] (\d+) tests, (\d+) expectations, (\d+) failed", result.stdout)
    assert summary, result.stdout[-2000:] + result.stderr[-2000:]
    tests, expectations, failed = (int(v) for v in summary.groups())
    assert result.returncode == 0 and failed == 0, result.stdout[-3000:]
    assert tests == EXPECTED[name] and expectations > 0
