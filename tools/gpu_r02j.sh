#!/bin/bash
# round 2, call j: tree with SoA coordinates + single final sort, L-expansion flags in the downward pass: tests, bench lines, M2L workloads
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_r02j.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], d.get("m2l") and round(d["m2l"]["frac_of_dfma_peak"],3))'
run() { echo "== $*"; timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q"; }
( run --npart-side 256
  run --npart-side 512
  run --npart-side 256 --nside 128
  run --npart-side 256 --nside 128 --disp-rms 2.0
  run --ic merger
  run --npart-side 512 --precision fp64 ) 2>&1 | tee gpurun_out/bench_quick_r02j.log
