#!/bin/bash
# Round-2 (second session) multi-GPU evidence with the final kernels (under gpurun --gpus N): the
# bench line at N ranks - sharded == one-GPU / reference parity keys, the strong-scaling 1024^3
# key and, at N = 8, the 2048^3 grid of BASELINE config 5 - and tests/multi_gpu_check.py.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N="${1:-8}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port 29511 bench.py --gpus "$N" --steps 20 --warmup 3 \
    2> gpurun_out/r2b_bench_n${N}.err | tail -1 > gpurun_out/r2b_bench_n${N}.json
tail -c 600 gpurun_out/r2b_bench_n${N}.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port 29513 tests/multi_gpu_check.py > gpurun_out/r2b_multi_gpu_check_w${N}.log 2>&1
tail -3 gpurun_out/r2b_multi_gpu_check_w${N}.log
python -c "
import json
j = json.load(open('gpurun_out/r2b_bench_n${N}.json'))
print(j['value'], j['ms_per_step'], j['roofline'].get('rank0_stage_ms'), j.get('parity'), j.get('strong_scaling_1024'), j.get('config5_2048'))"
