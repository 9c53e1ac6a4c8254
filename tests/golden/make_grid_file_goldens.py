#!/usr/bin/env python3
"""Writes tests/golden/grid_files/*: small SDFR / SDFZ / CMGR / CMGZ files produced by the
REFERENCE'S OWN SaveToFile members (SignedDistanceField<float/double>, OccupancyMap), compiled
unmodified into oracle/_ref (needs /root/reference at build time: `make -C oracle ref`). The
inputs are fixed by this script, so the files can be regenerated and compared. Run from the repo
root:  python tests/golden/make_grid_file_goldens.py"""
import ctypes
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
from oracle import reference_oracle  # noqa: E402

OUT = Path(__file__).resolve().parent / "grid_files"


def fixed_inputs():
    shape = (3, 2, 4)
    values = (np.arange(np.prod(shape), dtype=np.float64).reshape(shape) - 7.5) * 0.125
    cells = np.array([0.0, 0.5, 1.0, 0.25], dtype=np.float32)[np.arange(np.prod(shape)) % 4]
    cells = cells.reshape(shape)
    # a rotation by 90 degrees about z and a translation (exact in binary), row-major 4x4
    origin = np.array([[0.0, -1.0, 0.0, 1.5], [1.0, 0.0, 0.0, -2.25], [0.0, 0.0, 1.0, 0.5],
                       [0.0, 0.0, 0.0, 1.0]])
    return shape, values, cells, origin


def main():
    sdf_lib = ctypes.CDLL(str(reference_oracle._PATH))
    map_lib = ctypes.CDLL(str(reference_oracle._MAPS_PATH))
    shape, values, cells, origin = fixed_inputs()
    column_major = np.ascontiguousarray(origin.T.reshape(-1))
    OUT.mkdir(exist_ok=True)
    message = ctypes.create_string_buffer(256)
    index = {"shape": list(shape), "resolution": 0.25, "frame": "golden_frame", "oob_value": 9.5,
             "origin_row_major": origin.reshape(-1).tolist(), "files": {}}
    for dtype, tag in ((np.float32, "f32"), (np.float64, "f64")):
        data = np.ascontiguousarray(values.astype(dtype))
        for compress in (0, 1):
            for locked in (0, 1):
                name = f"sdf_{tag}_{'z' if compress else 'r'}_{'locked' if locked else 'open'}.bin"
                code = sdf_lib.vgt_ref_sdf_save_to_file(
                    ctypes.c_int(data.dtype.itemsize), data.ctypes.data_as(ctypes.c_void_p),
                    *(ctypes.c_int64(v) for v in shape), ctypes.c_double(0.25),
                    column_major.ctypes.data_as(ctypes.c_void_p), b"golden_frame",
                    ctypes.c_int(locked), ctypes.c_double(9.5), str(OUT / name).encode(),
                    ctypes.c_int(compress), message, ctypes.c_int64(256))
                assert code == 0, message.value
                index["files"][name] = {"kind": "sdf", "dtype": tag, "compressed": bool(compress),
                                        "locked": bool(locked)}
    for compress in (0, 1):
        name = f"map_{'z' if compress else 'r'}.bin"
        code = map_lib.vgt_ref_map_save_to_file(
            cells.ctypes.data_as(ctypes.c_void_p), *(ctypes.c_int64(v) for v in shape),
            ctypes.c_double(0.25), column_major.ctypes.data_as(ctypes.c_void_p), b"golden_frame",
            ctypes.c_float(0.5), ctypes.c_float(0.75), str(OUT / name).encode(),
            ctypes.c_int(compress), message, ctypes.c_int64(256))
        assert code == 0, message.value
        index["files"][name] = {"kind": "map", "compressed": bool(compress),
                                "default_occupancy": 0.5, "oob_occupancy": 0.75}
    (OUT / "index.json").write_text(json.dumps(index, indent=1) + "\n")
    print("wrote", len(index["files"]), "files into", OUT)


if __name__ == "__main__":
    main()
