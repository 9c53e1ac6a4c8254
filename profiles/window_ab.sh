#!/bin/bash
# A/B of the window kernel against the stack kernel on one B200 (run under gpurun).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests/test_gpu_window_kernel.py -x -q 2>&1 | tail -3
for n in ${SIZES:-512}; do
  echo "lean:"; VGT_B200_ENVELOPE=lean python profiles/time_passes.py $n 10
  echo "window, pilot:"; python profiles/time_passes.py $n 10
  echo "window, no pilot:"; VGT_B200_WINDOW_PILOT=0 python profiles/time_passes.py $n 10
done
python bench.py --steps 5 --warmup 3 --skip-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench', d['ms_per_step'], d['roofline']['pass_ms'])
print(' other', d.get('other_configs'), '| voxel sdf ms', d['voxelizer'].get('sdf_of_voxelized_map_ms'))
print(' maps', {k:v for k,v in d.get('other_map_types',{}).items() if 'ms' in k})
"
