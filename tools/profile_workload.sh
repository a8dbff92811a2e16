#!/bin/bash
# Full ncu capture of ONE later-axis launch on a structured workload (run on the GPU box, ONE GPU):
#   gpurun --timeout 600 -- 'bash tools/profile_workload.sh r02_voronoi voronoi 512 [skip]'
# perf_matrix.py runs 2 warm-up transforms and 1 timed one; launches of the tile kernel come in
# (Y, Z) pairs, so skip=4 profiles the Y pass and skip=5 the Z pass of the third transform.
# Leaves  gpurun_out/<tag>.ncu-rep, <tag>_ncu.json (summary), <tag>_source.csv (SASS + CUDA lines).
set -u
tag=${1:-wl}; wl=${2:-voronoi}; size=${3:-512}; skip=${4:-4}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:later_axis_tile_kernel -s $skip -c 1 -f \
    -o gpurun_out/${tag} python tools/perf_matrix.py --size $size --only $wl --steps 1 > gpurun_out/${tag}_run.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}.ncu-rep gpurun_out/${tag}_ncu.json
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_source.csv 2>/dev/null
ls -la gpurun_out/${tag}*
