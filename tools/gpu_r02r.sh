#!/bin/bash
# round 2, call r (1 GPU): chunked first pass (leaf kernel behind the frontier launches on a second stream): parity tests, chunk / resident-CTA sweep at 512^3,
# kernel-parameter variants (prebuilt, PN2GPU_LIB) at 256^3
mkdir -p gpurun_out
T=${1:-r02r}
export PYTHONFAULTHANDLER=1
( timeout -s ABRT 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) | tee gpurun_out/pytest_gpu_$T.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "lane_eff %.3f"%d["tiles"]["lane_efficiency_rank0"])'
run() { echo "== $*"; timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q"; }
( echo "## chunks 16 fctas 2 (default)"; run --npart-side 512
  for cf in "0 2" "8 2" "32 2" "16 1" "16 3" "16 4" "64 2"; do set -- $cf; echo "## PN2_WALK_CHUNKS=$1 PN2_WALK_FCTAS=$2"; PN2_WALK_CHUNKS=$1 PN2_WALK_FCTAS=$2 run --npart-side 512; done
  run --npart-side 256
  PN2_WALK_CHUNKS=0 run --npart-side 256
  for v in nb10 nb12 lb4 lb6; do echo "## variant $v"; PN2GPU_LIB=$PWD/photons-2.0_b200/variants/libpn2gpu_$v.so run --npart-side 256; done
  run --npart-side 256 --ic poisson
  run --npart-side 256 --precision fp64 ) 2>&1 | tee gpurun_out/bench_sweep_$T.log
