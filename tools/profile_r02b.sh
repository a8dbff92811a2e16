#!/bin/bash
# Re-capture of the structured / 1024^3 launches after the in-place stores (ONE GPU; summaries only):
#   gpurun --timeout 900 -- 'bash tools/profile_r02b.sh'
set -u
for spec in "cfg3 512 4 y" "cfg3 512 5 z" "cfg2b 512 5 z" "balls 512 5 z" "voronoi 512 4 y" "voronoi 512 5 z" \
            "cfg2 1024 4 y" "cfg2 1024 5 z" "cfg2b 1024 4 y" "cfg2b 1024 5 z" "cfg3 1024 5 z"; do
  set -- $spec
  bash tools/profile_workload.sh r02b_${1}_${2}_${4} $1 $2 $3 > /dev/null 2>&1
  rm -f gpurun_out/r02b_${1}_${2}_${4}.ncu-rep gpurun_out/r02b_${1}_${2}_${4}_source.csv gpurun_out/r02b_${1}_${2}_${4}_run.log
done
ncu --set full --clock-control none -k regex:first_axis_vec_kernel -s 2 -c 1 -f -o gpurun_out/r02b_cfg2_1024_x \
    python tools/perf_matrix.py --size 1024 --only cfg2 --steps 1 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r02b_cfg2_1024_x.ncu-rep gpurun_out/r02b_cfg2_1024_x_ncu.json
rm -f gpurun_out/r02b_cfg2_1024_x.ncu-rep
ls gpurun_out/r02b_*_ncu.json
