#!/bin/bash
# round 2, call n (1 GPU): tests; ragged kernel on / off for the sparse-leaf workloads; compute-sanitizer; default bench line + reference arm
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_r02n.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "lane_eff %.3f"%d["tiles"]["lane_efficiency_rank0"])'
run() { echo "== $*"; timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q"; }
( PN2_RAGGED=0 run --npart-side 256 --ic poisson
  PN2_RAGGED=1 run --npart-side 256 --ic poisson
  PN2_RAGGED=0 run --npart-side 256 --disp-rms 2.0
  PN2_RAGGED=1 run --npart-side 256 --disp-rms 2.0
  PN2_RAGGED=1 run --npart-side 256
  PN2_RAGGED=0 run --npart-side 256 ) 2>&1 | tee gpurun_out/sweep_ragged_r02n.log
bash tools/gpu_sanitize.sh r02n
python bench.py > gpurun_out/bench_512_r02n.json 2> gpurun_out/bench_512_r02n.err; tail -c 1500 gpurun_out/bench_512_r02n.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02n.json 2> gpurun_out/bench_ref_r02n.err; tail -c 300 gpurun_out/bench_ref_r02n.json
