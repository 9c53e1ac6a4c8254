"""CPU: the C-ABI library builds, loads and exports every symbol include/vgt_b200.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

from voxelized_geometry_tools_b200 import _capi

REPO = Path(__file__).resolve().parents[1]
HEADER = REPO / "include" / "vgt_b200.h"


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"VGT_B200_API\s+[\w\s\*]+?\b(vgt_b200_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    assert len(names) >= 16
    for needed in ("vgt_b200_sdf_f32", "vgt_b200_sdf_f64", "vgt_b200_sdf_from_mask_f32",
                   "vgt_b200_edt_sq_i32", "vgt_b200_voxelize_f64", "vgt_b200_raycast_f64_dev",
                   "vgt_b200_filter_dev", "vgt_b200_edt_local_passes_dev",
                   "vgt_b200_edt_final_pass_f32_dev", "vgt_b200_last_error"):
        assert needed in names


def test_library_exports_every_declared_symbol(shared_library):
    handle = ctypes.CDLL(str(shared_library))
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} is declared in the header but not exported"
    exported = subprocess.run(["nm", "-D", "--defined-only", str(shared_library)],
                              capture_output=True, text=True).stdout
    public = sorted(set(re.findall(r" T (vgt_b200_\w+)", exported)))
    assert public == declared_symbols(), "exported C symbols and header declarations differ"


def test_python_binding_table_matches_header(shared_library):
    assert sorted(_capi.SIGNATURES) == declared_symbols()
    lib = _capi.library()
    assert lib.vgt_b200_version().startswith(b"vgt_b200")
    assert lib.vgt_b200_last_error() is not None


def test_every_entry_point_cites_the_reference():
    # Each declaration block names the reference interface it replaces (file:line).
    text = HEADER.read_text()
    blocks = re.split(r"\n(?=/\*)", text)
    cited = [b for b in blocks if "VGT_B200_API int vgt_b200_" in b]
    assert cited
    for block in cited:
        if "profile" in block:
            continue
        assert re.search(r"\.(hpp|cpp|cu):\d+", block), block[:200]


def test_no_cpu_fallback_without_device(shared_library):
    # In the CPU-only container the device count is 0 and every host-side entry refuses to run.
    import voxelized_geometry_tools_b200 as vgt
    import numpy as np
    if _capi.device_count() > 0:
        pytest.skip("a GPU is present")
    sizes = vgt.VoxelGridSizes.FromVoxelCounts(0.1, (4, 4, 4))
    occupancy_map = vgt.OccupancyMap(np.eye(4), "f", sizes)
    with pytest.raises(vgt.BackendUnavailable):
        occupancy_map.ExtractSignedDistanceFieldFloat(vgt.SignedDistanceFieldGenerationParameters())
    with pytest.raises(vgt.BackendUnavailable):
        vgt.B200PointCloudVoxelizer()
    assert vgt.GetAvailableBackends() == []
    # and the raw C call reports a device error, not a silent success
    out = np.zeros((4, 4, 4), dtype=np.float32)
    code = _capi.library().vgt_b200_sdf_f32(occupancy_map.GetImmutableRawData().ctypes.data, 4, 4,
                                            4, 0.1, 1, 0, 0, out.ctypes.data, None, None)
    assert code == _capi.ERR_DEVICE
    assert "Cuda error" in _capi.last_error()
