#!/bin/bash
# End-of-session evidence run on one B200 (under gpurun): GPU tests, bench lines, launch list,
# full ncu capture of the three SDF kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/r1g_bench_n1.json 2> gpurun_out/r1g_bench_n1.err
tail -c 600 gpurun_out/r1g_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1g_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:Kernel$ -c 600 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/r1g_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ScanContiguous|EnvelopeAxis' -c 6 -o gpurun_out/r1g_kernels python profiles/run_sdf_once.py 512 1 > gpurun_out/r1g_kernels.log 2>&1
tail -2 gpurun_out/r1g_kernels.log
cat gpurun_out/r1g_bench_n1.json | cut -c1-1500
