#!/bin/bash
# Window kernel: blocks per multiprocessor the line segmenting aims at (VGT_B200_WINDOW_BLOCKS_PER_SM).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for b in ${BLOCKS:-48 96 192 288 384}; do
  echo "== blocks per SM $b"
  for n in ${SIZES:-512}; do VGT_B200_WINDOW_BLOCKS_PER_SM=$b python profiles/time_passes.py $n 10; done
done
