"""ctypes front-end for oracle/_ref/libvgt_ref.so (TEST INFRASTRUCTURE ONLY).

That library is the REFERENCE'S OWN signed_distance_field_generation.{hpp,cpp}, compiled
unmodified from /root/reference over the stand-in headers in oracle/ref_shim/ (``make -C oracle
ref``). It is used (a) to pin the restated oracle (tests/test_oracle_vs_reference.py) and (b) as
the CPU baseline of kind "reference" in bench.py. It is built in the dev container and travels
to the GPU box prebuilt (oracle/_ref/ is git-ignored but not gpurun-ignored).
"""
from __future__ import annotations

import ctypes
from pathlib import Path

import numpy as np

_PATH = Path(__file__).resolve().parent / "_ref" / "libvgt_ref.so"
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64
_int = ctypes.c_int
_lib = None


def available() -> bool:
    return _PATH.exists()


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        handle = ctypes.CDLL(str(_PATH))
        handle.vgt_ref_sdf_f32.argtypes = [_f32p, _i64, _i64, _i64, ctypes.c_double, _int, _int,
                                           _int, _f32p, _f32p]
        handle.vgt_ref_sdf_f64.argtypes = [_f32p, _i64, _i64, _i64, ctypes.c_double, _int, _int,
                                           _int, _f64p, _f64p]
        handle.vgt_ref_transform_inplace_f64.argtypes = [_f64p, _i64, _i64, _i64, _int]
        _lib = handle
    return _lib


def sdf(occupancy, resolution: float, unknown_is_filled: bool = True,
        add_virtual_border: bool = False, threads: int = 0, dtype=np.float32):
    occ = np.ascontiguousarray(occupancy, dtype=np.float32)
    out = np.empty(occ.shape, dtype=dtype)
    min_max = np.zeros(2, dtype=dtype)
    if np.dtype(dtype) == np.float32:
        code = lib().vgt_ref_sdf_f32(
            occ.ctypes.data_as(_f32p), *occ.shape, float(resolution), int(unknown_is_filled),
            int(add_virtual_border), threads, out.ctypes.data_as(_f32p),
            min_max.ctypes.data_as(_f32p))
    else:
        code = lib().vgt_ref_sdf_f64(
            occ.ctypes.data_as(_f32p), *occ.shape, float(resolution), int(unknown_is_filled),
            int(add_virtual_border), threads, out.ctypes.data_as(_f64p),
            min_max.ctypes.data_as(_f64p))
    if code != 0:
        raise RuntimeError("reference SDF generation failed")
    return out, (min_max[0], min_max[1])


def transform_inplace(field: np.ndarray, threads: int = 0) -> np.ndarray:
    assert field.dtype == np.float64 and field.flags.c_contiguous and field.ndim == 3
    if lib().vgt_ref_transform_inplace_f64(field.ctypes.data_as(_f64p), *field.shape, threads):
        raise RuntimeError("reference transform failed")
    return field
