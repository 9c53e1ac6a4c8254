#!/bin/bash
# Turns the gpurun_out/r2b_* artefacts of profiles/r2b_run.sh / r2b_run_multi.sh into the committed
# summaries under profiles/ (run here, where ncu can read the reports).
cd "$(dirname "$0")/.."
python profiles/ncu_summary.py gpurun_out/r2b_kernels.ncu-rep > profiles/r2b_ncu_kernels.txt
python profiles/ncu_summary.py gpurun_out/r2b_voxelizer_kernels.ncu-rep > profiles/r2b_ncu_voxelizer.txt
python profiles/ncu_traffic.py gpurun_out/r2b_kernels.ncu-rep 512
cp gpurun_out/r2b_launches.csv profiles/r2b_launches.csv
for f in r2b_bench_n1.json r2b_bench_reference.json r2b_bench_n2.json r2b_bench_n4.json r2b_bench_n8.json \
         r2b_config5_8gpu.jsonl r2b_multi_gpu_check_w2.log r2b_multi_gpu_check_w8.log \
         r2b_multi_entry_n2.json r2b_multi_entry_n8.json r2b_gpu_tests.txt r2b_slab_times.json; do
  [ -f gpurun_out/$f ] && cp gpurun_out/$f profiles/$f
done
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(l for l in open("profiles/r2b_launches.csv") if l.startswith('"')))
header = rows[0]
name_i, value_i = header.index("Kernel Name"), header.index("Metric Value")
unit_i = header.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}
total = collections.Counter(); count = collections.Counter()
for row in rows[1:]:
    name = re.sub(r"\(.*", "", row[name_i]).replace("void ", "").replace("vgt_b200::", "")
    name = re.sub(r"<unnamed>::|unnamed>::|\(anonymous namespace\)::", "", name)
    ms = float(row[value_i].replace(",", "")) * scale.get(row[unit_i], 1e-6)
    total[name] += ms; count[name] += 1
whole = sum(total.values())
with open("profiles/r2b_launches.md", "w") as out:
    out.write("# ncu launch list of `bench.py --steps 3 --warmup 3 --skip-cpu --skip-strong` (round 2, second session: final kernels)\n\n")
    out.write("`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are "
              "cold-cache and serialised: the SHARE of each kernel is what matters.\n\n")
    out.write("| kernel | launches | total ms | share | mean us |\n|---|---|---|---|---|\n")
    for name, ms in total.most_common():
        out.write(f"| `{name[:100]}` | {count[name]} | {ms:.3f} | {100*ms/whole:.1f} % | {1e3*ms/count[name]:.1f} |\n")
print(open("profiles/r2b_launches.md").read()[:3000])
PY
