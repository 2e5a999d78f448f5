#!/bin/bash
# round 2, call u (1 GPU): 128-entry span fetch in the leaf walk and the frontier kernel: parity tests, bench lines; frontier kernel at 6 / 7 CTAs per SM
mkdir -p gpurun_out
T=${1:-r02u}
export PYTHONFAULTHANDLER=1
( timeout -s ABRT 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) | tee gpurun_out/pytest_gpu_$T.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "step %.3f"%d["roofline"]["whole_step"]["frac"], "lane_eff %.3f"%d["tiles"]["lane_efficiency_rank0"])'
run() { echo "== $*"; timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q"; }
( run --npart-side 256
  run --npart-side 512
  for v in nb6 nb7; do echo "## variant $v"; PN2GPU_LIB=$PWD/photons-2.0_b200/variants/libpn2gpu_$v.so run --npart-side 256; done
  run --npart-side 256 --ic poisson
  run --npart-side 256 --precision fp64 ) 2>&1 | tee gpurun_out/bench_quick_$T.log
