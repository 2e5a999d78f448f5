#!/bin/bash
# round 2, call ab (1 GPU): ncu --set full of the M2L kernel (256^3, NSIDE 128: 1.07e7 pairs) -- what bounds it
mkdir -p gpurun_out
T=${1:-r02ab}
CMD="python bench.py --npart-side 256 --nside 128 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:m2l_warp_kernel -s 1 -c 1 -o gpurun_out/prof_m2l_$T -f $CMD > gpurun_out/prof_m2l_$T.log 2>&1
ls -la gpurun_out/prof_m2l_$T.ncu-rep
