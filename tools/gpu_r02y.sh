#!/bin/bash
# round 2, call y (1 GPU): last check of the judged code: all parity tests, smoke(), bench lines (uniform / clustered: tree phase with the bounded level look-ahead)
mkdir -p gpurun_out
T=${1:-r02y}
export PYTHONFAULTHANDLER=1
( timeout -s ABRT 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) | tee gpurun_out/pytest_gpu_$T.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) | tee gpurun_out/smoke_$T.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({"cmd": sys.argv[1], "ms": round(d["ms_per_step"],2), "pps": d["value"], "phases": {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac": round(d["roofline"]["frac"],3), "step_frac": round(d["roofline"]["whole_step"]["frac"],3)}))'
run() { timeout 400 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q" "$*"; }
( run --npart-side 512
  run --npart-side 256 --disp-rms 2.0
  run --ic merger
  run --npart-side 32 ) 2>&1 | tee gpurun_out/bench_check_$T.jsonl
