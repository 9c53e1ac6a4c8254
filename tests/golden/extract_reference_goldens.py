#!/usr/bin/env python3
"""Extracts the known-answer values of the reference's own gtest files into JSON fixtures.

Run in the dev container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/extract_reference_goldens.py

It READS the reference test sources and writes small fixtures next to itself:

* ``sdf_generation_test.json``  <- test/sdf_generation_test.cpp
  per TEST_P case: grid extents + resolution, the filled box (loop bounds), the
  default occupancy, expected (min, max) with the test's own tolerance, and every
  ``EXPECT_FLOAT_EQ(get_occupancy_map_sdf_dist(x, y, z), value)`` cell.
* ``pointcloud_voxelization_test.json`` <- test/pointcloud_voxelization_test.cpp
  the scene constants the test hard-codes (grid, camera poses, cloud lattice, filter
  options); the plane-level expectations are restated in tests/test_voxelizer_oracle.py
  citing the same lines.
* ``voxel_raycasting_test.json`` <- test/voxel_raycasting_test.cpp (grid, range, seed).

Nothing here is copied code: only literals (numbers) are lifted.
"""
from __future__ import annotations

import json
import math
import re
from pathlib import Path

REFERENCE = Path("/root/reference")
HERE = Path(__file__).resolve().parent


def _literal(text: str) -> float:
    text = text.strip()
    match = re.fullmatch(r"std::sqrt\(([-0-9.]+)f?\)", text)
    if match:
        # The test evaluates std::sqrt(2.0f) in float.
        import numpy as np
        return float(np.sqrt(np.float32(float(match.group(1)))))
    return float(text.rstrip("f"))


def extract_sdf_generation() -> dict:
    source = (REFERENCE / "test" / "sdf_generation_test.cpp").read_text()
    lines = source.splitlines()
    tolerance = float(re.search(r"kExtremaTolerance = ([0-9.]+);", source).group(1))
    starts = [i for i, line in enumerate(lines) if line.startswith("TEST_P(")]
    starts.append(len(lines))
    cases = []
    for begin, end in zip(starts[:-1], starts[1:]):
        body = lines[begin:end]
        text = "\n".join(body)
        name = re.match(r"TEST_P\(\w+, (\w+)\)", body[0]).group(1)
        case = {
            "name": name,
            "reference_lines": [begin + 1, end],
            "resolution": float(re.search(r"resolution = ([0-9.]+);", text).group(1)),
            "size_xyz": [float(re.search(rf"{axis}_size = ([0-9.]+);", text).group(1))
                         for axis in "xyz"],
        }
        default_cell = re.search(r"grid_sizes, OccupancyCell\(([0-9.]+)f\)\)", text)
        case["default_occupancy"] = float(default_cell.group(1))
        # Filled region: explicit loop bounds, or a whole z = const layer.
        box = {}
        for axis in "xyz":
            loop = re.search(
                rf"for \(int64_t {axis}_index = (\d+); {axis}_index < (\w+);", text)
            if loop:
                upper = loop.group(2)
                box[axis] = [int(loop.group(1)), int(upper) if upper.isdigit() else None]
        layer = re.search(r"constexpr int64_t z_index = (\d+);", text)
        if layer:
            box["z"] = [int(layer.group(1)), int(layer.group(1)) + 1]
        case["filled_box"] = box if ("SetIndex" in text) else None
        # Expected extrema (float variant).
        if "negative_inf" in text:
            case["expected_min_max"] = ["-inf", "-inf"]
        elif "positive_inf" in text:
            case["expected_min_max"] = ["inf", "inf"]
        else:
            mn = re.search(r"const float minimum = ([-0-9.]+)f;", text)
            mx = re.search(r"const float maximum = ([-0-9.]+)f;", text)
            if mn and mx:
                case["expected_min_max"] = [float(mn.group(1)), float(mx.group(1))]
            elif mn:
                nominal = re.search(
                    r"std::sqrt\(std::pow\(resolution, 2.0\) \+\s*"
                    r"std::pow\(([0-9.]+) \* resolution, 2.0\) \+\s*"
                    r"std::pow\(([0-9.]+)\s*\* resolution, 2.0\)\)", text)
                r = case["resolution"]
                maximum = math.sqrt(r ** 2 + (float(nominal.group(1)) * r) ** 2
                                    + (float(nominal.group(2)) * r) ** 2)
                case["expected_min_max"] = [float(mn.group(1)), maximum]
        cells = []
        for i, line in enumerate(body):
            if "EXPECT_FLOAT_EQ(get_occupancy_map_sdf_dist(" in line:
                joined = line.strip()
                if not joined.endswith(";"):
                    joined += " " + body[i + 1].strip()
                match = re.match(
                    r"EXPECT_FLOAT_EQ\(get_occupancy_map_sdf_dist\((\d+), (\d+), (\d+)\), (.+)\);",
                    joined)
                cells.append({"index": [int(match.group(k)) for k in (1, 2, 3)],
                              "value": _literal(match.group(4)),
                              "line": begin + i + 1})
        case["expected_cells"] = cells
        cases.append(case)
    return {"source": "test/sdf_generation_test.cpp",
            "extrema_tolerance": tolerance,
            "generation_parameters": {"oob_value": "inf", "unknown_is_filled": True,
                                      "add_virtual_border": False},
            "cases": cases}


def extract_voxelization() -> dict:
    source = (REFERENCE / "test" / "pointcloud_voxelization_test.cpp").read_text()

    def number(pattern: str) -> float:
        return float(re.search(pattern, source).group(1))

    lattice = re.search(
        r"for \(double x = ([-0-9.]+); x <= ([-0-9.]+); x \+= ([0-9.]+)\)", source)
    return {
        "source": "test/pointcloud_voxelization_test.cpp",
        "grid_origin_translation": [float(v) for v in re.search(
            r"X_WG\(Eigen::Translation3d\(([-0-9.]+), ([-0-9.]+), ([-0-9.]+)\)\)",
            source).groups()],
        "grid_size_xyz": [number(rf"const double {a}_size = ([0-9.]+);") for a in "xyz"],
        "grid_resolution": number(r"grid_resolution = ([0-9.]+);"),
        "lattice_min_max_step": [float(v) for v in lattice.groups()],
        "near_depth": number(r"\? ([0-9.]+) : [0-9.]+;"),
        "far_depth": number(r"\? [0-9.]+ : ([0-9.]+);"),
        "camera1_translation": [float(v) for v in re.search(
            r"X_WC1\(Eigen::Translation3d\(([-0-9.]+), ([-0-9.]+), ([-0-9.]+)\)\)",
            source).groups()],
        "camera2_translation": [float(v) for v in re.search(
            r"X_WC2 = Eigen::Translation3d\(([-0-9.]+), ([-0-9.]+), ([-0-9.]+)\)",
            source).groups()],
        "percent_seen_free": number(r"percent_seen_free = ([0-9.]+);"),
        "outlier_points_threshold": int(number(r"outlier_points_threshold = (\d+);")),
        "num_cameras_seen_free": int(number(r"num_cameras_seen_free = (\d+);")),
    }


def extract_raycasting() -> dict:
    source = (REFERENCE / "test" / "voxel_raycasting_test.cpp").read_text()
    return {
        "source": "test/voxel_raycasting_test.cpp",
        "resolution": float(re.search(r"resolution = ([0-9.]+);", source).group(1)),
        "voxel_counts": [int(v) for v in re.search(
            r"Vector3i64\((\d+), (\d+), (\d+)\)", source).groups()],
        "min_axis_value": float(re.search(r"min_axis_value = ([-0-9.]+);", source).group(1)),
        "max_axis_value": float(re.search(r"max_axis_value = ([-0-9.]+);", source).group(1)),
        "max_range": float(re.search(r"max_range = ([0-9.]+);", source).group(1)),
        "seed": int(re.search(r"std::mt19937_64 prng\((\d+)\)", source).group(1)),
        "iterations": int(re.search(r"iterations = (\d+);", source).group(1)),
    }


def main() -> None:
    for name, payload in (("sdf_generation_test.json", extract_sdf_generation()),
                          ("pointcloud_voxelization_test.json", extract_voxelization()),
                          ("voxel_raycasting_test.json", extract_raycasting())):
        (HERE / name).write_text(json.dumps(payload, indent=1) + "\n")
        print("wrote", HERE / name)


if __name__ == "__main__":
    main()
