#!/bin/bash
# round 2, call aa (1 GPU): software-pipelined M2L kernel (cp.async operand tiles, two in flight): parity tests, M2L workloads, occupancy variants
mkdir -p gpurun_out
T=${1:-r02aa}
export PYTHONFAULTHANDLER=1
( timeout -s ABRT 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) | tee gpurun_out/pytest_gpu_$T.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({"cmd": sys.argv[1], "ms": round(d["ms_per_step"],2), "m2l_ms": round(d["phases_ms"]["m2l"],2), "m2l": {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.get("m2l",{}).items() if k!="kernel"}}))'
run() { timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q" "$*"; }
( run --npart-side 256 --nside 128
  run --npart-side 256 --nside 128 --disp-rms 2.0
  run --ic merger
  for v in m2lmb4 m2lmb8; do echo "## variant $v"; PN2GPU_LIB=$PWD/photons-2.0_b200/variants/libpn2gpu_$v.so run --npart-side 256 --nside 128; PN2GPU_LIB=$PWD/photons-2.0_b200/variants/libpn2gpu_$v.so run --ic merger; done ) 2>&1 | tee gpurun_out/bench_m2l_$T.jsonl
