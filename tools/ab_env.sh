#!/bin/bash
# Same-box comparison of (library, environment) pairs:
#   tools/ab_env.sh "libA.so:VAR=1 libB.so:" [workloads] [size] [steps]
# (each pair timed twice, interleaved, so that box-to-box and warm-up effects show)
pairs=$1; wl=${2:-cfg2,cfg2b,cfg3,balls,voronoi}; size=${3:-512}; steps=${4:-10}
for round in 1 2; do
  for pair in $pairs; do
    lib=${pair%%:*}; envs=${pair#*:}
    echo "== $lib [$envs] (round $round)"
    env EDTB200_LIBRARY=$PWD/$lib ${envs//,/ } python tools/perf_matrix.py --size $size --only $wl --steps $steps | cut -c1-100
  done
done
