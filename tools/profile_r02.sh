#!/bin/bash
# Round-2 profiling batch (ONE GPU): bench launch list + full captures of the headline kernels,
# full captures on the structured workloads, per-axis captures at 1024^3 (BASELINE configs[3]),
# compute-sanitizer on the all-variant case list.  Summaries only come back (reports are dropped).
set -u
bash tools/profile_round.sh r02_final
for spec in "cfg3 512 5 z" "cfg2 1024 4 y" "cfg2 1024 5 z" "cfg2b 1024 4 y" "cfg2b 1024 5 z"; do
  set -- $spec
  bash tools/profile_workload.sh r02_${1}_${2}_${4} $1 $2 $3 > /dev/null 2>&1
  rm -f gpurun_out/r02_${1}_${2}_${4}.ncu-rep gpurun_out/r02_${1}_${2}_${4}_source.csv
done
# the X pass at 1024^3
ncu --set full --clock-control none -k regex:first_axis_vec_kernel -s 2 -c 1 -f -o gpurun_out/r02_cfg2_1024_x \
    python tools/perf_matrix.py --size 1024 --only cfg2 --steps 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02_cfg2_1024_x.ncu-rep gpurun_out/r02_cfg2_1024_x_ncu.json
rm -f gpurun_out/r02_cfg2_1024_x.ncu-rep
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool python tools/sanitize_cases.py > gpurun_out/r02_sanitizer_$tool.txt 2>&1
  tail -4 gpurun_out/r02_sanitizer_$tool.txt
done
ls gpurun_out/r02_*
