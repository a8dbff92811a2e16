#!/bin/bash
# Round-end profiling recipe (run on the GPU box through gpurun, ONE GPU):
#   gpurun --timeout 600 -- 'bash tools/profile_round.sh r01_final'
# 1. launch list of the bench command (gpu__time_duration.sum, no clock control) -> gpurun_out/<tag>_launches.csv
# 2. one `--set full` capture of the later-axis tile kernel (the Z pass is the dominant kernel; the
#    first two profiled launches are the Y and Z passes of one transform) -> gpurun_out/<tag>_tile.ncu-rep
# 3. the same for the first-axis kernel -> gpurun_out/<tag>_first.ncu-rep
# Numbers printed by bench.py under ncu are NOT bench values.
set -u
tag=${1:-round}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1
# (launches of the tile kernel come in Y, Z pairs; the first transforms of a process still use the
#  3-CTA variant -- the 2-CTA one is chosen once a transform has published its run statistic -- so
#  skip well into the warm-up: launches 10 and 11 are the Y and Z pass of the sixth transform)
ncu --set full --clock-control none --import-source on -k regex:later_axis_tile_kernel -s 10 -c 2 -f \
    -o gpurun_out/${tag}_tile python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:first_axis_vec_kernel -s 4 -c 1 -f \
    -o gpurun_out/${tag}_first python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_tile.ncu-rep gpurun_out/${tag}_later_axis_ncu.json
python tools/ncu_summary.py gpurun_out/${tag}_first.ncu-rep gpurun_out/${tag}_first_axis_ncu.json
# gpurun brings back at most 64 MiB: keep the summaries, drop the reports unless asked to keep them
if [ "${KEEP_REPORTS:-0}" != "1" ]; then rm -f gpurun_out/${tag}_tile.ncu-rep gpurun_out/${tag}_first.ncu-rep; fi
ls -la gpurun_out/ | head -20
