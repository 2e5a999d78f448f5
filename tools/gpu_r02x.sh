#!/bin/bash
# round 2, call x (1 GPU): FP64 kernel with the degree-6 / 31-interval table of g(u): parity tests, smoke(), FP64 bench lines; FP32 check line
mkdir -p gpurun_out
T=${1:-r02x}
export PYTHONFAULTHANDLER=1
( timeout -s ABRT 500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "128\^3|demo nside 32 precision 0|passed|failed|Error|error" | tail -12 ) | tee gpurun_out/pytest_gpu_$T.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 ) | tee gpurun_out/smoke_$T.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({"cmd": sys.argv[1], "ms": round(d["ms_per_step"],2), "pps": d["value"], "phases": {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac": round(d["roofline"]["frac"],3), "step_frac": round(d["roofline"]["whole_step"]["frac"],3)}))'
run() { timeout 400 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q" "$*"; }
( run --npart-side 256 --precision fp64
  run --npart-side 512 --precision fp64
  run --npart-side 512 ) 2>&1 | tee gpurun_out/bench_fp64_$T.jsonl
