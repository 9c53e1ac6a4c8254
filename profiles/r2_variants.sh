#!/bin/bash
# A/B of build-time variants of the window kernel on one B200 (under gpurun). The variants are
# built beforehand (nvcc -D...) into voxelized_geometry_tools_b200/_variants/NAME.so; each one is
# copied over the library of this scratch copy, checked for parity (the SDF must not change) and
# timed pass by pass.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P=voxelized_geometry_tools_b200
cp $P/libvgt_b200.so /tmp/libvgt_b200_shipped.so
for v in ${VARIANTS}; do
  cp $P/_variants/$v.so $P/libvgt_b200.so
  echo "== $v"
  if [ -n "$PARITY" ]; then python -m pytest tests/test_gpu_sdf.py tests/test_gpu_window_kernel.py -x -q -m gpu 2>&1 | tail -1; fi
  for n in ${SIZES:-512}; do python profiles/time_passes.py $n 10; done
done
cp /tmp/libvgt_b200_shipped.so $P/libvgt_b200.so
