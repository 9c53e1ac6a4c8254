#!/bin/bash
# A/B of the window kernel against the stack kernel on one B200 (run under gpurun).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests/test_gpu_window_kernel.py -x -q 2>&1 | tail -5
for n in ${SIZES:-512 256}; do
  echo "lean:"; VGT_B200_ENVELOPE=lean python profiles/time_passes.py $n 10
  echo "window default:"; python profiles/time_passes.py $n 10
  echo "window 6 blocks:"; VGT_B200_WINDOW_BLOCKS=6 python profiles/time_passes.py $n 10
  echo "window 8 blocks:"; VGT_B200_WINDOW_BLOCKS=8 python profiles/time_passes.py $n 10
done
