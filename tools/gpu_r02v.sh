#!/bin/bash
# round 2, call v (1 GPU): the state to be judged: all parity tests (verbose numbers), the default bench line + reference arm, per-mode and
# workload lines, ncu launch list + full captures of the two walk kernels, compute-sanitizer
mkdir -p gpurun_out
T=${1:-r02v}
export PYTHONFAULTHANDLER=1
( timeout -s ABRT 600 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "rms rel|err |passed|failed|Error|error|skipped|density|pair" | tail -80 ) | tee gpurun_out/pytest_gpu_$T.log
timeout 600 python bench.py > gpurun_out/bench_512_$T.json 2> gpurun_out/bench_512_$T.err; tail -c 2500 gpurun_out/bench_512_$T.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; tail -c 400 gpurun_out/bench_ref_$T.json
Q='import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({"cmd": sys.argv[1], "ms": round(d["ms_per_step"],2), "pps": d["value"], "phases": {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac": round(d["roofline"]["frac"],3), "step_frac": round(d["roofline"]["whole_step"]["frac"],3), "lane_eff": round(d["tiles"]["lane_efficiency_rank0"],3), "m2l": {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.get("m2l",{}).items() if k!="kernel"}}))'
run() { timeout 400 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q" "$*"; }
( run --npart-side 512 --precision fp64
  run --npart-side 256
  run --npart-side 256 --precision fp64
  run --npart-side 256 --ic poisson
  run --npart-side 256 --disp-rms 2.0
  run --npart-side 256 --maxleaf 16
  run --npart-side 256 --maxleaf 32
  run --npart-side 256 --nside 128
  run --npart-side 256 --nside 128 --disp-rms 2.0
  run --ic merger
  run --npart-side 32
  run --npart-side 64 ) 2>&1 | tee gpurun_out/bench_workloads_$T.jsonl
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/launches_$T.log 2>&1
CMD="python bench.py --npart-side 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_fused_kernel -s 1 -c 1 -o gpurun_out/prof_fused_$T -f $CMD > gpurun_out/prof_fused_$T.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frontier_node -s 44 -c 3 -o gpurun_out/prof_frontier_$T -f $CMD > gpurun_out/prof_frontier_$T.log 2>&1
bash tools/gpu_sanitize.sh $T
ls gpurun_out | grep $T | tr '\n' ' '
