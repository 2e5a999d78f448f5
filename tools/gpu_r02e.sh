#!/bin/bash
# round 2, call e: M2L warp kernel + 2^27 cell ids + vectorised top levels: parity tests; workload-spread bench lines
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_r02e.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "lane_eff %.3f"%d["tiles"]["lane_efficiency_rank0"], d.get("m2l"))'
for s in 256 512; do timeout 300 python bench.py --npart-side $s --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "$Q"; done 2>&1 | tee gpurun_out/bench_quick_r02e.log
: > gpurun_out/bench_workloads_r02e.jsonl
run() { echo "== $*"; timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>gpurun_out/last.err | tail -1 | tee -a gpurun_out/bench_workloads_r02e.jsonl | python -c "$Q"; }
( run --npart-side 256 --disp-rms 2.0
  run --npart-side 256 --ic poisson
  run --npart-side 256 --maxleaf 16
  run --npart-side 256 --maxleaf 32
  run --npart-side 256 --nside 128
  run --npart-side 256 --nside 128 --disp-rms 2.0
  run --npart-side 256 --nside 128 --precision fp64
  run --ic merger
  run --ic merger --precision fp64 ) 2>&1 | tee gpurun_out/bench_workloads_r02e.log
