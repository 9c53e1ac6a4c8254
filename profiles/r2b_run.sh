#!/bin/bash
# Round-2 (second session, final kernels) evidence run on ONE B200 (under gpurun): GPU tests, bench lines (ours + reference arm),
# ncu launch list of the bench command, `ncu --set full` of the SDF kernels and of the voxelizer
# kernels. Everything lands in gpurun_out/r2b_*; profiles/r2b_collect.sh turns the reports into the
# text summaries that are committed.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2b_gpu_tests.txt
cat gpurun_out/r2b_gpu_tests.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
tail -c 400 gpurun_out/r2b_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'Kernel$' -c 900 --csv \
    --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 3 --warmup 3 --skip-cpu --skip-strong > gpurun_out/r2b_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ScanContiguous|EnvelopeAxis' -c 8 \
    -o gpurun_out/r2b_kernels python profiles/run_sdf_once.py 512 1 > gpurun_out/r2b_kernels.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'RaycastCloud|FilterGrids' -c 5 \
    -o gpurun_out/r2b_voxelizer_kernels python profiles/run_voxelizer_once.py > gpurun_out/r2b_voxelizer_kernels.log 2>&1
tail -2 gpurun_out/r2b_kernels.log gpurun_out/r2b_voxelizer_kernels.log
cut -c1-1200 gpurun_out/r2b_bench_n1.json
