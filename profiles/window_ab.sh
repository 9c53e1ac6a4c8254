#!/bin/bash
# A/B of the window kernel variants on one B200 (run under gpurun).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests/test_gpu_window_kernel.py -x -q 2>&1 | tail -3
for n in ${SIZES:-512}; do
  echo "window (register prefetch):"; python profiles/time_passes.py $n 10
  echo "window (cp.async staged):"; VGT_B200_WINDOW_STAGE=1 python profiles/time_passes.py $n 10
done
