#!/bin/bash
# Same-box comparison of several builds of the library:
#   tools/ab.sh "build/libA.so build/libB.so ..." [workloads] [size]
# (each library timed twice, interleaved, so that box-to-box and warm-up effects show)
libs=$1; wl=${2:-cfg2,cfg2b,cfg3,balls,voronoi}; size=${3:-512}
for round in 1 2; do
  for lib in $libs; do
    echo "== $lib (round $round)"
    EDTB200_LIBRARY=$PWD/$lib python tools/perf_matrix.py --size $size --only $wl --steps 10 | cut -c1-110
  done
done
