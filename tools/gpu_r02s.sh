#!/bin/bash
# round 2, call s (1 GPU): table-driven M2L kernel (no libm), leaf-kernel launch refactor: parity tests; M2L workloads; R degree 7 variant with the parity block
mkdir -p gpurun_out
T=${1:-r02s}
export PYTHONFAULTHANDLER=1
( timeout -s ABRT 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) | tee gpurun_out/pytest_gpu_$T.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "lane_eff %.3f"%d["tiles"]["lane_efficiency_rank0"], {k:(round(v,4) if isinstance(v,float) else v) for k,v in d.get("m2l",{}).items() if k!="kernel"}, "parity", d.get("parity",{}).get("fp32_rms"), d.get("parity",{}).get("fp64_rms"))'
run() { echo "== $*"; timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q"; }
( run --npart-side 256 --no-cpu-baseline --nside 128
  run --npart-side 256 --no-cpu-baseline --nside 128 --disp-rms 2.0
  run --no-cpu-baseline --ic merger
  run --npart-side 256 --no-cpu-baseline
  run --npart-side 512 --no-cpu-baseline
  echo "## R degree 8 (default) with the parity block"; run --npart-side 256
  echo "## R degree 7 with the parity block"; PN2GPU_LIB=$PWD/photons-2.0_b200/variants/libpn2gpu_rdeg7.so run --npart-side 256
  PN2GPU_LIB=$PWD/photons-2.0_b200/variants/libpn2gpu_rdeg7.so run --npart-side 512 --no-cpu-baseline ) 2>&1 | tee gpurun_out/bench_workloads_$T.log
