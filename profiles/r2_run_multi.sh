#!/bin/bash
# Round-2 multi-GPU evidence (under gpurun --gpus N): the bench line at N ranks (with the
# sharded == single-GPU / oracle parity keys), tests/multi_gpu_check.py at world N, BASELINE
# config 5 at N = 8 and the single-process multi-device entry.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N="${1:-8}"
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port 29511 bench.py --gpus "$N" --steps 20 --warmup 3 \
    2> gpurun_out/r2_bench_n${N}.err | tail -1 > gpurun_out/r2_bench_n${N}.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port 29513 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_w${N}.log 2>&1
tail -3 gpurun_out/r2_multi_gpu_check_w${N}.log
if [ "$N" = "8" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29515 profiles/run_config5.py 2>/dev/null | grep '^{' > gpurun_out/r2_config5_8gpu.jsonl
  cat gpurun_out/r2_config5_8gpu.jsonl
fi
python profiles/r2_multi_entry.py 1024 > gpurun_out/r2_multi_entry_n${N}.json 2>&1
cat gpurun_out/r2_multi_entry_n${N}.json
python -c "
import json
j = json.load(open('gpurun_out/r2_bench_n${N}.json'))
print(j['value'], j['ms_per_step'], j['roofline']['rank0_stage_ms'], j.get('parity'), j.get('strong_scaling_1024'))"
