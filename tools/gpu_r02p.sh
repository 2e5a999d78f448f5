#!/bin/bash
# round 2, call p (1 GPU): localise the 64^3 hang of call o (every command under its own short timeout), ncu capture of the new fused kernel
mkdir -p gpurun_out
T=${1:-r02p}
export PYTHONFAULTHANDLER=1
( timeout -s ABRT 200 python -m pytest tests/test_gpu_mode_b.py -m gpu -x -q 2>&1 | tail -15 ) | tee gpurun_out/dbg_modeb_$T.log
( timeout -s ABRT 200 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -15 ) | tee gpurun_out/dbg_multirank_$T.log
( timeout -s ABRT 120 python bench.py --npart-side 64 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -c 1500 ) | tee gpurun_out/dbg_bench64_e2e_$T.log
( timeout -s ABRT 240 python bench.py --npart-side 64 --steps 2 --warmup 3 --cpu-sample-side 32 2>&1 | tail -c 3000 ) | tee gpurun_out/dbg_bench64_full_$T.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "lane_eff %.3f"%d["tiles"]["lane_efficiency_rank0"])'
run() { echo "== $*"; timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>gpurun_out/last.err | tail -1 | python -c "$Q"; }
( run --npart-side 256 ) 2>&1 | tee gpurun_out/bench_quick_$T.log
CMD="python bench.py --npart-side 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_fused_kernel -s 1 -c 1 -o gpurun_out/prof_fused_$T -f $CMD > gpurun_out/prof_fused_$T.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frontier_node -s 44 -c 3 -o gpurun_out/prof_frontier_$T -f $CMD > gpurun_out/prof_frontier_$T.log 2>&1
ls -la gpurun_out | tail -4
