#!/bin/bash
# Last evidence run of the third session on one B200 (under gpurun): GPU tests and the bench line
# with the final window-kernel configuration (one-warp blocks, depth bound max(64, length / 8)).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r1h_bench_n1.json 2> gpurun_out/r1h_bench_n1.err
tail -c 300 gpurun_out/r1h_bench_n1.err
cut -c1-400 gpurun_out/r1h_bench_n1.json
