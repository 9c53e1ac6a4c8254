"""CPU: the SASS of the built library carries the instructions the design relies on (cuobjdump
reads the sm_100a cubin without a GPU). A regression here means the compiler no longer emits what
DESIGN.md section 3 describes, whatever the timings say later:
  * the window kernels' candidate search is VIADDMNMX.U16x2 (two rows per instruction), the
    nearest opposite-class row comes from FLO + VIMNMX3 + SHFL, and the bookkeeping around them
    (addresses, shifts, bit extraction) is multiplies on the fma pipe;
  * the staged window kernels fetch their rows with LDGSTS (cp.async) and wait on LDGDEPBAR /
    DEPBAR groups, the others with plain loads;
  * the z scan moves 16 bytes per lane and instruction;
  * the voxelizer's counters are RED atomics (no return value) and its arithmetic keeps separate
    double multiplies and adds (-fmad=false: the DDA must reproduce the CPU's double-precision
    steps; the DFMAs that remain belong to the correctly rounded division sequences)."""
import re
import shutil
import subprocess

import pytest


@pytest.fixture(scope="module")
def sass_by_function(shared_library):
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    result = subprocess.run([cuobjdump, "-sass", str(shared_library)], capture_output=True,
                            text=True, timeout=300)
    assert result.returncode == 0, result.stderr[-2000:]
    functions, name = {}, None
    for line in result.stdout.splitlines():
        match = re.search(r"Function : (\S+)", line)
        if match:
            name = match.group(1)
            functions[name] = []
        elif name is not None:
            functions[name].append(line)
    assert functions, "no device functions found in the library"
    return {key: "\n".join(lines) for key, lines in functions.items()}


def functions_named(sass_by_function, fragment):
    found = {k: v for k, v in sass_by_function.items() if fragment in k}
    assert found, f"no kernel named *{fragment}* in the library"
    return found


def test_library_is_sm_100a_only(shared_library):
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    listing = subprocess.run([cuobjdump, "-lelf", str(shared_library)], capture_output=True,
                             text=True, timeout=120).stdout
    archs = set(re.findall(r"sm_\d+a?", listing))
    assert archs == {"sm_100a"}, archs


def test_window_kernels_use_the_packed_add_min(sass_by_function):
    kernels = functions_named(sass_by_function, "EnvelopeAxisWindowKernel")
    for name, sass in kernels.items():
        radius = int(re.search(r"EnvelopeAxisWindowKernelILi\dELi(\d+)E", name).group(1))
        packed = sass.count("VIADDMNMX.U16x2")
        # three unrolled chunk bodies (interior with / without the class arithmetic, edge) x R
        # rows x (R + 1) pairs, plus the chunk-joint search (register tier + four rows per side
        # and round from memory, R / 2 paired add-mins each)
        assert packed >= 3 * radius * (radius + 1) + 8 * (radius // 2), (name, packed)
        # nearest opposite-class row: boundary word reversed once per chunk, one count of
        # leading zeros per side, a three-way minimum and a look-up across the warp
        assert "BREV" in sass and "FLO" in sass, name
        assert "SHFL.IDX" in sass and "VIMNMX3" in sass, name
        # row addresses, bit extraction and shifts by constants run on the fma pipe: multiplies
        # by powers of two read from constant bank 3 (ptxas cannot fold them back into shifts)
        assert len(re.findall(r"IMAD(\.HI)?(\.U32)? R\d+, R\d+, (c\[0x3\]|UR\d+)", sass)) >= 4 * radius, name


def test_staged_window_kernels_use_async_copies(sass_by_function):
    # kStage (the last template argument): 0 = register prefetch, 1 = per-lane cp.async
    # (LDGSTS), 2 = cp.async.bulk through the TMA engine (UBLKCP) with mbarrier completion (SYNCS)
    kernels = functions_named(sass_by_function, "EnvelopeAxisWindowKernel")
    by_stage = {0: [], 1: [], 2: []}
    for name in kernels:
        stage = int(re.search(r"ELi\d+ELi(\d)EEEv", name).group(1))
        by_stage[stage].append(name)
    assert by_stage[0] and by_stage[1] and by_stage[2], {k: len(v) for k, v in by_stage.items()}
    for name in by_stage[0]:
        assert "LDGSTS" not in kernels[name] and "UBLKCP" not in kernels[name], name
    for name in by_stage[1]:
        assert "LDGSTS" in kernels[name], name
        assert "LDGDEPBAR" in kernels[name] or "DEPBAR" in kernels[name], name
        assert "UBLKCP" not in kernels[name], name
    for name in by_stage[2]:
        assert "UBLKCP" in kernels[name], name          # the bulk copies
        assert "SYNCS" in kernels[name], name           # mbarrier init / arrive / try_wait
        assert "LDGSTS" not in kernels[name], name


def test_z_scan_moves_sixteen_bytes_per_lane(sass_by_function):
    for name, sass in functions_named(sass_by_function, "ScanContiguousAxisRegistersKernel").items():
        if "Float4Source" in name or "12Float4Source" in name:
            assert re.search(r"LDG\.E(\.[A-Z]+)*\.128", sass), name
        assert re.search(r"STG\.E(\.[A-Z]+)*\.128", sass), name


def test_voxelizer_uses_reductions_and_separate_multiplies(sass_by_function):
    raycast = functions_named(sass_by_function, "RaycastCloudKernel")
    for name, sass in raycast.items():
        assert re.search(r"\bRED(G)?\.", sass), name
        assert "ATOMG" not in sass, name              # no atomic that returns a value
        assert sass.count("DMUL") >= 10 and sass.count("DADD") >= 10, name
