#!/bin/bash
# A/B of the exchange variants at N ranks (under gpurun --gpus N): parity check first, then the
# bench line (main timing + stage times) per variant.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N="${1:-2}"
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port 29513 tests/multi_gpu_check.py > gpurun_out/r2b_multi_gpu_check_w${N}.log 2>&1
tail -4 gpurun_out/r2b_multi_gpu_check_w${N}.log
for variant in "peer_store 1" "peer_copy 2" "peer_copy 4" "peer_copy 8"; do
  set -- $variant
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus "$N" --steps 20 --warmup 3 --exchange $1 --chunks $2 \
      --skip-strong --skip-config5 --skip-parity 2> gpurun_out/r2b_ab.err | tail -1 > gpurun_out/r2b_ab_${1}_${2}_n${N}.json
  python -c "
import json
j = json.load(open('gpurun_out/r2b_ab_${1}_${2}_n${N}.json'))
print('$1 chunks $2:', round(j['value'], 1), 'Gvoxels/s', round(j['ms_per_step'], 4), 'ms', j['roofline'].get('rank0_stage_ms'), j['roofline'].get('exchange'))" || tail -5 gpurun_out/r2b_ab.err
done
