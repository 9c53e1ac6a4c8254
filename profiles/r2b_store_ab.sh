#!/bin/bash
# A/B of build variants on the sharded path at N ranks (under gpurun --gpus N).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N="${1:-2}"
P=voxelized_geometry_tools_b200
cp $P/libvgt_b200.so /tmp/libvgt_b200_shipped.so
for v in shipped ${VARIANTS}; do
  if [ "$v" = shipped ]; then cp /tmp/libvgt_b200_shipped.so $P/libvgt_b200.so; else cp $P/_variants/$v.so $P/libvgt_b200.so; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus "$N" --steps 20 --warmup 3 \
      --skip-strong --skip-config5 ${EXTRA} 2> gpurun_out/r2b_ab.err | tail -1 > gpurun_out/r2b_ab_${v}_n${N}.json
  python -c "
import json
j = json.load(open('gpurun_out/r2b_ab_${v}_n${N}.json'))
print('$v:', round(j['value'], 1), 'Gvoxels/s', round(j['ms_per_step'], 4), 'ms', j['roofline'].get('rank0_stage_ms'), j.get('parity'))" || tail -5 gpurun_out/r2b_ab.err
done
cp /tmp/libvgt_b200_shipped.so $P/libvgt_b200.so
