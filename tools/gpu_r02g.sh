#!/bin/bash
# round 2, call g: 30-bit Morton pre-sort, top target 1024: parity tests, bench lines, ncu captures (fused FP32, frontier twig level)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_r02g.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"])'
for s in 256 512; do timeout 300 python bench.py --npart-side $s --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "$Q"; done 2>&1 | tee gpurun_out/bench_quick_r02g.log
CMD="python bench.py --npart-side 256 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_fused_kernel -s 1 -c 1 -o gpurun_out/prof_fused_r02g -f $CMD > gpurun_out/prof_fused_r02g.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frontier_node -s 44 -c 3 -o gpurun_out/prof_frontier_r02g -f $CMD > gpurun_out/prof_frontier_r02g.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
