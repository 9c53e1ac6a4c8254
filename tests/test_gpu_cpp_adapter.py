"""GPU: runs the prebuilt C++ adapter test (voxelized_geometry_tools_b200/cpp/test/adapter_test.cpp),
which ports the reference's gtests through the C++ host adapter compiled against the reference's
own pointcloud_voxelization_interface.hpp. The binary is built in the dev container
(`make -C voxelized_geometry_tools_b200/cpp`, also done by __graft_entry__.build()) because
/root/reference does not exist on the GPU box."""
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

BINARY = (Path(__file__).resolve().parents[1] / "voxelized_geometry_tools_b200" / "cpp" / "_build"
          / "adapter_test")


def test_cpp_adapter_ports_of_the_reference_tests(shared_library):
    if not BINARY.exists():
        pytest.skip("adapter_test was not prebuilt (needs /root/reference at build time)")
    result = subprocess.run([str(BINARY)], capture_output=True, text=True, timeout=300)
    assert result.returncode == 0, result.stdout + result.stderr
    assert "ADAPTER_TEST_OK" in result.stdout
