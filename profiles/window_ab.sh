#!/bin/bash
# A/B of the window kernel against the stack kernel on one B200 (run under gpurun).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests/test_gpu_window_kernel.py -x -q 2>&1 | tail -3
for n in ${SIZES:-512}; do
  echo "lean:"; VGT_B200_ENVELOPE=lean python profiles/time_passes.py $n 10
  for sg in 4 5 6 8; do
    for t in 0 1 2 3; do echo "segments $sg tune $t:"; VGT_B200_WINDOW_SEGMENTS=$sg VGT_B200_WINDOW_TUNE=$t python profiles/time_passes.py $n 10; done
  done
  for b in 7; do echo "window $b blocks 4 seg:"; VGT_B200_WINDOW_SEGMENTS=4 VGT_B200_WINDOW_BLOCKS=$b python profiles/time_passes.py $n 10; done
done
