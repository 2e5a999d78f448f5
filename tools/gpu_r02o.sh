#!/bin/bash
# round 2, call o (1 GPU): leaf walk that queues source leaves untested (nodes only on the stack): parity tests + quick bench lines
mkdir -p gpurun_out
T=${1:-r02o}
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_$T.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "lane_eff %.3f"%d["tiles"]["lane_efficiency_rank0"])'
run() { echo "== $*"; timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q"; }
( run --npart-side 256
  run --npart-side 512
  run --npart-side 256 --ic poisson
  run --npart-side 256 --precision fp64 ) 2>&1 | tee gpurun_out/bench_quick_$T.log
