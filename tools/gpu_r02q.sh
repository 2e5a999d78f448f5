#!/bin/bash
# round 2, call q (1 GPU): leaf walk v2 (tile ids in the queue, leaf header inside the tile, node-only stack), flattened frontier step,
# bounded grids for the pass over the received trees: all parity tests, quick bench lines, ncu capture of the fused kernel
mkdir -p gpurun_out
T=${1:-r02q}
export PYTHONFAULTHANDLER=1
( timeout -s ABRT 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) | tee gpurun_out/pytest_gpu_$T.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "lane_eff %.3f"%d["tiles"]["lane_efficiency_rank0"])'
run() { echo "== $*"; timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q"; }
( run --npart-side 256
  run --npart-side 512
  run --npart-side 256 --ic poisson
  run --npart-side 256 --precision fp64 ) 2>&1 | tee gpurun_out/bench_quick_$T.log
CMD="python bench.py --npart-side 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_fused_kernel -s 1 -c 1 -o gpurun_out/prof_fused_$T -f $CMD > gpurun_out/prof_fused_$T.log 2>&1
ls -la gpurun_out | tail -3
