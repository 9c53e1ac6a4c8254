#!/bin/bash
# A/B of the window kernel against the stack kernel on one B200 (run under gpurun).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests/test_gpu_window_kernel.py -x -q 2>&1 | tail -3
for n in ${SIZES:-512 256}; do
  echo "lean:"; VGT_B200_ENVELOPE=lean python profiles/time_passes.py $n 10
  echo "window, pilot:"; python profiles/time_passes.py $n 10
  echo "window, no pilot:"; VGT_B200_WINDOW_PILOT=0 python profiles/time_passes.py $n 10
done
