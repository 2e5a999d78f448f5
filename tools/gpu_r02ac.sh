#!/bin/bash
# round 2, call ac (1 GPU): M2L kernel at 4 and 2 lanes per sink for short lists: M2L parity tests, NSIDE = n/2 workload per width, clustered, merger
mkdir -p gpurun_out
T=${1:-r02ac}
export PYTHONFAULTHANDLER=1
( timeout -s ABRT 300 python -m pytest tests/test_gpu_mode_b.py tests/test_gpu_mode_a.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3 ) | tee gpurun_out/pytest_gpu_$T.log
( PN2_M2L_LPS=2 timeout -s ABRT 200 python -m pytest tests/test_gpu_mode_b.py -m gpu -x -q 2>&1 | tail -2 ) | tee -a gpurun_out/pytest_gpu_$T.log
( PN2_M2L_LPS=4 timeout -s ABRT 200 python -m pytest tests/test_gpu_mode_b.py -m gpu -x -q 2>&1 | tail -2 ) | tee -a gpurun_out/pytest_gpu_$T.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({"cmd": sys.argv[1], "ms": round(d["ms_per_step"],2), "m2l_ms": round(d["phases_ms"]["m2l"],2), "m2l": {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.get("m2l",{}).items() if k!="kernel"}}))'
run() { timeout 200 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q" "$*"; }
( for w in 2 4 8; do echo "## PN2_M2L_LPS=$w"; PN2_M2L_LPS=$w run --npart-side 256 --nside 128; done
  echo "## automatic"; run --npart-side 256 --nside 128
  run --ic merger ) 2>&1 | tee gpurun_out/bench_m2l_$T.jsonl
