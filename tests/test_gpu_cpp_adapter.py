"""GPU: runs the prebuilt C++ adapter test (voxelized_geometry_tools_b200/cpp/test/adapter_test.cpp),
which ports the reference's gtests through the C++ host adapter compiled against the reference's
own pointcloud_voxelization_interface.hpp. The binary is built in the dev container
(`make -C voxelized_geometry_tools_b200/cpp`, also done by __graft_entry__.build()) because
/root/reference does not exist on the GPU box."""
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

BUILD = Path(__file__).resolve().parents[1] / "voxelized_geometry_tools_b200" / "cpp" / "_build"


@pytest.mark.parametrize("name", ["adapter_test", "adapter_test_fast_lock"])
def test_cpp_adapter_ports_of_the_reference_tests(shared_library, name):
    # adapter_test_fast_lock = the same with the proposed SignedDistanceField accessor compiled in
    binary = BUILD / name
    if not binary.exists():
        pytest.skip(f"{name} was not prebuilt (needs /root/reference at build time)")
    result = subprocess.run([str(binary)], capture_output=True, text=True, timeout=300)
    assert result.returncode == 0, result.stdout + result.stderr
    assert "ADAPTER_TEST_OK" in result.stdout


def test_cpp_adapter_timing_leg(shared_library):
    import json
    binary = BUILD / "adapter_test"
    if not binary.exists():
        pytest.skip("adapter_test was not prebuilt (needs /root/reference at build time)")
    result = subprocess.run([str(binary), "--time", "128"], capture_output=True, text=True,
                            timeout=300)
    assert result.returncode == 0, result.stdout + result.stderr
    line = json.loads(result.stdout.strip().splitlines()[-1])
    assert line["grid"] == "128^3" and line["ms_per_call"] > 0.0
