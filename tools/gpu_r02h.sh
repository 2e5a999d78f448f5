#!/bin/bash
# round 2, call h: tests; tree switch-depth sweep at 512^3; FP64 interval index variants at 256^3
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_r02h.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"])'
for t in 1024 256 64 16; do echo "== PN2_TREE_TOP_TARGET=$t"; PN2_TREE_TOP_TARGET=$t timeout 300 python bench.py --npart-side 512 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "$Q"; done 2>&1 | tee gpurun_out/sweep_tree_r02h.log
for v in "-DF64_MAGIC=1" "-DF64_MAGIC=0"; do
  touch photons-2.0_b200/csrc/pn2_walk.cu
  make -s -C photons-2.0_b200/csrc EXTRA="$v" > /dev/null 2>&1
  echo "== EXTRA=$v"
  timeout 300 python bench.py --precision fp64 --npart-side 256 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "$Q"
done 2>&1 | tee gpurun_out/sweep_f64_r02h.log
