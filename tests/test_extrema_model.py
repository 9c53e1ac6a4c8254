"""CPU: a numpy model of the parallel local-extrema-map algorithm of csrc/sdf_queries.cu
(successors -> pointer jumping -> smallest basin cell per loop -> loop entry -> write) against the
reference's sequential memoising loop as restated in oracle/sdf_queries_oracle.py. Checks the
argument the kernels rest on: the sequential result depends on the order of the walks only
through the cell where the first walk of a basin enters its loop."""
import numpy as np
import pytest

from oracle import sdf_queries_oracle

OFF_GRID = -1


def parallel_model(checker):
    dims = (checker.nx, checker.ny, checker.nz)
    count = int(np.prod(dims))
    successor = np.empty(count, dtype=np.int64)
    for cell, index in enumerate(np.ndindex(*dims)):
        g = checker.coarse_gradient_at_index(*index, True)[1]
        if checker._effectively_flat(g):
            successor[cell] = cell
            continue
        moved = checker._next_from_gradient(index, g)
        successor[cell] = (np.ravel_multi_index(moved, dims) if checker._in_bounds(moved)
                           else OFF_GRID)
    target, lowest = successor.copy(), np.arange(count)
    rounds = 1
    while (1 << rounds) < count:
        rounds += 1
    for _ in range(rounds + 1):
        on_grid = target != OFF_GRID
        safe = np.where(on_grid, target, 0)
        lowest = np.where(on_grid, np.minimum(lowest, lowest[safe]), lowest)
        target = np.where(on_grid, target[safe], target)
    on_grid = target != OFF_GRID
    safe = np.where(on_grid, target, 0)
    in_loop_basin = on_grid & (successor[safe] != safe)
    loop_id = np.where(in_loop_basin, lowest[safe], -1)
    entry = {}
    for loop in np.unique(loop_id[in_loop_basin]):
        first = int(np.flatnonzero(loop_id == loop)[0])
        seen, current = {first}, first
        while True:
            current = int(successor[current])
            if current in seen:
                entry[int(loop)] = current
                break
            seen.add(current)
    out = np.full((count, 3), np.inf)
    for cell in range(count):
        if not on_grid[cell]:
            continue
        end = entry[int(loop_id[cell])] if in_loop_basin[cell] else int(target[cell])
        out[cell] = checker._centre_in_grid_frame(np.unravel_index(end, dims))
    return out.reshape(dims + (3,)), len(entry)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_parallel_extrema_model_equals_the_sequential_loop(seed):
    rng = np.random.default_rng(seed)
    dims = (9, 8, 10)
    field = rng.normal(size=dims).astype(np.float32) * 0.3
    field[rng.random(dims) < 0.1] = 0.0
    if seed == 3:
        field[2:5, 2:5, 2:5] = np.inf
    checker = sdf_queries_oracle.SdfOracle(field, 0.1, np.eye(4))
    got, loops = parallel_model(checker)
    assert loops > 3
    np.testing.assert_array_equal(got, checker.local_extrema_map())
