#!/usr/bin/env python3
"""Times vgt_b200_sdf_f32_multi (one process, all visible devices, host buffers in and out)
against the one-device host entry on the same grid, and checks that the results are identical.
    python profiles/r2_multi_entry.py [n]      (default 512: a 512 x 512 x 512 grid)"""
import ctypes
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from voxelized_geometry_tools_b200 import _capi, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = _capi.library()
count = _capi.device_count()
occupancy = synthetic.clustered_spheres_occupancy((n, n, n))
out_single = np.zeros_like(occupancy)
out_multi = np.zeros_like(occupancy)
lo, hi = ctypes.c_float(), ctypes.c_float()


def timed(call, reps=3):
    call()
    times = []
    for _ in range(reps):
        begin = time.perf_counter()
        call()
        times.append(time.perf_counter() - begin)
    return min(times) * 1e3


single_ms = timed(lambda: _capi.check(lib.vgt_b200_sdf_f32(
    occupancy.ctypes.data, n, n, n, 0.02, 1, 0, 0, out_single.ctypes.data, ctypes.byref(lo),
    ctypes.byref(hi))))
single_extrema = (lo.value, hi.value)
results = {"grid": f"{n}^3", "devices_visible": count, "one_device_ms": single_ms, "multi": []}
for used in sorted({2, 4, count} & set(range(2, count + 1))):
    listed = (ctypes.c_int * used)(*range(used))
    multi_ms = timed(lambda: _capi.check(lib.vgt_b200_sdf_f32_multi(
        occupancy.ctypes.data, n, n, n, 0.02, 1, 0, listed, used, out_multi.ctypes.data,
        ctypes.byref(lo), ctypes.byref(hi))))
    results["multi"].append({"devices": used, "ms": multi_ms,
                             "equals_one_device": bool(np.array_equal(out_multi, out_single))
                             and (lo.value, hi.value) == single_extrema})
print(json.dumps(results))
