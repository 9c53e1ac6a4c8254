#!/usr/bin/env python3
"""Round-2 experiment (one GPU): what each rank of the slab-sharded SDF would compute, timed
without any exchange, plus the depth distribution the window kernels see.

    python profiles/r2_slab_times.py [--out gpurun_out/r2_slab_times.json]

For the N = 2 / 4 / 8 bench grids it times, on ONE device,
  * the slab-local passes (z scan + y envelope, send layout) of every x-slab,
  * the final pass (x envelope + finalize) of every y-slab,
so that compute imbalance between ranks can be told apart from NVLink effects in the sharded
runs. It also histograms the distances (in voxels) after the y pass and after the x pass: the
share of voxels deeper than the register window R decides how much extended search runs.
"""
from __future__ import annotations

import argparse
import json
import statistics
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import torch  # noqa: E402

from voxelized_geometry_tools_b200 import device as vdev, sharded, synthetic  # noqa: E402

RESOLUTION = 0.02


def timed(fn, reps=3):
    times = []
    for _ in range(reps):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        fn()
        stop.record()
        torch.cuda.synchronize()
        times.append(start.elapsed_time(stop))
    return statistics.median(times)


def depth_histogram(squared: torch.Tensor):
    """Share of voxels whose distance (voxels) exceeds each bound; squared = int64/float tensor."""
    bounds = [4, 8, 12, 16, 24, 32, 48, 64, 96, 128]
    total = squared.numel()
    return {str(b): float((squared > b * b).sum().item()) / total for b in bounds}


def run(dims, world):
    dev = torch.device("cuda", 0)
    nx, ny, nz = dims
    occupancy = synthetic.clustered_spheres_occupancy_torch(dims, dev)
    out = torch.empty_like(occupancy)
    min_max = torch.empty(2, dtype=torch.float32, device=dev)
    for _ in range(2):
        vdev.signed_distance_field(occupancy, RESOLUTION, out=out, min_max=min_max)
    torch.cuda.synchronize()
    passes = [vdev.signed_distance_field_profile(occupancy, RESOLUTION, out, min_max)
              for _ in range(3)]
    whole = [statistics.median(p[i] for p in passes) for i in range(3)]
    result = {"dims": list(dims), "world": world, "one_gpu_pass_ms": whole,
              "one_gpu_total_ms": sum(whole)}
    final_depth = depth_histogram((out.abs() / RESOLUTION).square_())
    result["share_deeper_than_after_x"] = final_depth

    packed = vdev.edt_local_passes(occupancy)
    torch.cuda.synchronize()
    magnitudes = (packed & 0x7fffffff).to(torch.float32)
    result["share_deeper_than_after_y"] = depth_histogram(magnitudes)
    del magnitudes

    local_ms, final_ms = [], []
    for rank in range(world):
        x0, x1 = sharded.split_range(nx, world, rank)
        slab = occupancy[x0:x1]
        vdev.edt_local_passes(slab, send_parts=world)
        local_ms.append(timed(lambda: vdev.edt_local_passes(slab, send_parts=world)))
    for rank in range(world):
        y0, y1 = sharded.split_range(ny, world, rank)
        y_slab = packed[:, y0:y1, :].contiguous()
        work = torch.empty_like(y_slab)

        def final():
            work.copy_(y_slab)
            vdev.edt_final_pass(work, y0, ny, RESOLUTION)

        copy_ms = timed(lambda: work.copy_(y_slab))
        final()
        final_ms.append(timed(final) - copy_ms)
    result["local_passes_ms_per_rank"] = local_ms
    result["final_pass_ms_per_rank"] = final_ms
    result["ideal_ms"] = result["one_gpu_total_ms"] / world
    result["max_local_plus_max_final_ms"] = max(local_ms) + max(final_ms)
    return result


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--out", default=str(REPO / "gpurun_out" / "r2_slab_times.json"))
    parser.add_argument("--only", type=int, default=0)
    args = parser.parse_args()
    cases = [((512, 512, 512), 1), ((1024, 512, 512), 2), ((1024, 1024, 512), 4),
             ((1024, 1024, 1024), 8)]
    results = []
    for dims, world in cases:
        if args.only and world != args.only:
            continue
        results.append(run(dims, world))
        print(json.dumps(results[-1]), flush=True)
        torch.cuda.empty_cache()
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
