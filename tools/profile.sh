#!/bin/bash
# GPU-side profiling recipe (run under gpurun): launch list of the default bench workload + one full capture of
# the dominant kernel (and of the frontier kernel).  usage: tools/profile.sh <tag> [npart-side of the full capture]
TAG=${1:-r01}; SIDE=${2:-256}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/launches_$TAG.log 2>&1
CMD="python bench.py --npart-side $SIDE --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:walk_fused -s 1 -c 1 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/prof_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:frontier_node -s 40 -c 3 -o gpurun_out/prof_frontier_$TAG -f $CMD > gpurun_out/prof_frontier_$TAG.log 2>&1
ls -la gpurun_out | tail -8
