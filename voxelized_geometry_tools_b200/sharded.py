"""Slab-sharded signed distance field across the GPUs of one box (one process per GPU).

Scheme: the grid is cut into x-slabs -- x is the slowest storage axis, so a
slab is one contiguous range of the reference's host vector. The z and y passes are independent
per x and run slab-local. One all-to-all then re-cuts the sign-fused int32 intermediate into
y-slabs laid out [nx, ny_local, nz], on which the x pass + finalize run, again with z contiguous.
min/max is reduced with one 2-float all-reduce. The result stays y-sharded on the devices;
``gather_to_host`` writes the slabs back to their x-major host positions (tests, small grids).

The reference has no multi-device path at all; this is the B200-native
analogue of its OpenMP split over lines (sdfgen.cpp:286-389).

The compute stages are injectable so the exchange logic can be tested on CPU with the gloo
backend (tests/test_sharded_gloo.py); the product stages are the CUDA ones in ``device.py``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import torch
import torch.distributed as dist


def split_range(total: int, parts: int, index: int) -> tuple[int, int]:
    """[begin, end) of part ``index`` when ``total`` items are cut into ``parts`` near-equal parts
    (the first ``total % parts`` parts get one extra item)."""
    base, extra = divmod(total, parts)
    begin = index * base + min(index, extra)
    return begin, begin + base + (1 if index < extra else 0)


@dataclass
class Stages:
    """local_passes(occupancy_slab[nxl, ny, nz] f32, unknown_is_filled) -> int32 [nxl, ny, nz]
    final_pass(packed[nx, nyl, nz] int32, y_offset, ny_total, resolution, add_virtual_border)
        -> (sdf f32 [nx, nyl, nz], min_max f32 [2])"""
    local_passes: Callable
    final_pass: Callable


def cuda_stages() -> Stages:
    from . import device

    def local_passes(occupancy_slab, unknown_is_filled):
        return device.edt_local_passes(occupancy_slab, unknown_is_filled)

    def final_pass(packed, y_offset, ny_total, resolution, add_virtual_border):
        return device.edt_final_pass(packed, y_offset, ny_total, resolution, add_virtual_border)

    return Stages(local_passes, final_pass)


class ShardedSignedDistanceField:
    def __init__(self, dims, rank: int | None = None, world_size: int | None = None,
                 group=None, stages: Stages | None = None):
        self.dims = tuple(int(d) for d in dims)
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        nx, ny, _ = self.dims
        if self.world_size > min(nx, ny):
            raise ValueError("more ranks than voxels along x or y")
        self.stages = stages if stages is not None else cuda_stages()
        self.x_range = split_range(nx, self.world_size, self.rank)
        self.y_range = split_range(ny, self.world_size, self.rank)

    # ------------------------------------------------------------------ layout helpers
    def x_slab_shape(self):
        return (self.x_range[1] - self.x_range[0], self.dims[1], self.dims[2])

    def y_slab_shape(self):
        return (self.dims[0], self.y_range[1] - self.y_range[0], self.dims[2])

    def _exchange(self, packed: torch.Tensor) -> torch.Tensor:
        """All-to-all transpose: x-slab [nxl, ny, nz] -> y-slab [nx, nyl, nz].

        The block received from rank g is rows x in [x0_g, x1_g) of the y-slab, which is one
        contiguous range of the destination, so the collective writes straight into place."""
        nx, ny, nz = self.dims
        world = self.world_size
        nyl = self.y_range[1] - self.y_range[0]
        if world == 1:
            return packed
        send_chunks = []
        for peer in range(world):
            y0, y1 = split_range(ny, world, peer)
            send_chunks.append(packed[:, y0:y1, :].reshape(-1))
        send = torch.cat(send_chunks)
        send_splits = [int(chunk.numel()) for chunk in send_chunks]
        recv_splits = []
        for peer in range(world):
            x0, x1 = split_range(nx, world, peer)
            recv_splits.append((x1 - x0) * nyl * nz)
        received = torch.empty(nx * nyl * nz, dtype=packed.dtype, device=packed.device)
        dist.all_to_all_single(received, send, recv_splits, send_splits, group=self.group)
        return received.view(nx, nyl, nz)

    # ------------------------------------------------------------------ the path
    def extract(self, occupancy_slab: torch.Tensor, resolution: float,
                unknown_is_filled: bool = True, add_virtual_border: bool = False):
        """occupancy_slab: this rank's x-slab [nxl, ny, nz] float32.
        Returns (sdf y-slab [nx, nyl, nz] float32, global (min, max) tensor [2])."""
        if tuple(occupancy_slab.shape) != self.x_slab_shape():
            raise ValueError(f"expected an x-slab of shape {self.x_slab_shape()}")
        packed = self.stages.local_passes(occupancy_slab, unknown_is_filled)
        y_slab = self._exchange(packed)
        sdf, min_max = self.stages.final_pass(
            y_slab, self.y_range[0], self.dims[1], resolution, add_virtual_border)
        if self.world_size > 1:
            # one tiny all-reduce: max over (-min, max)
            folded = torch.stack([-min_max[0], min_max[1]])
            dist.all_reduce(folded, op=dist.ReduceOp.MAX, group=self.group)
            min_max = torch.stack([-folded[0], folded[1]])
        return sdf, min_max

    def gather_to_host(self, sdf_y_slab: torch.Tensor) -> torch.Tensor | None:
        """Collects the y-slabs into one [nx, ny, nz] host tensor on rank 0 (small grids only)."""
        nx, ny, nz = self.dims
        local = sdf_y_slab.detach().cpu().contiguous()
        if self.world_size == 1:
            return local
        pieces = [None] * self.world_size if self.rank == 0 else None
        dist.gather_object(local, pieces, dst=0, group=self.group)
        if self.rank != 0:
            return None
        full = torch.empty((nx, ny, nz), dtype=local.dtype)
        for peer, piece in enumerate(pieces):
            y0, y1 = split_range(ny, self.world_size, peer)
            full[:, y0:y1, :] = piece
        return full
