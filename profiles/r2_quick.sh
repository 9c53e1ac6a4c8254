#!/bin/bash
# quick GPU check of an envelope-kernel change: parity tests of the SDF path, then pass timings
set -x
timeout 600 python -m pytest tests/test_gpu_sdf.py tests/test_gpu_window_kernel.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python profiles/r2_slab_times.py --out gpurun_out/r2_slab_times_quick.json 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        r = json.loads(line)
        print(r['dims'], 'passes', [round(v, 3) for v in r['one_gpu_pass_ms']], 'total', round(r['one_gpu_total_ms'], 3),
              'local', [round(v, 2) for v in r['local_passes_ms_per_rank']], 'final', [round(v, 2) for v in r['final_pass_ms_per_rank']])
    else:
        print(line, end='')
"
