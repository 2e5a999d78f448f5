#!/bin/bash
# round 2, call b: FP64 table kernel: parity tests, FP64 bench lines (256^3, 512^3), ncu capture of the FP64 kernel
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_r02b.log
( timeout 600 python -m pytest tests/test_gpu_large_properties.py tests/test_gpu_mode_b.py -m gpu -x -q -s 2>&1 | grep -E "rms rel|passed|failed|Error" | tail -40 ) | tee gpurun_out/pytest_gpu_r02b_verbose.log
for s in 256 512; do
  timeout 900 python bench.py --precision fp64 --npart-side $s --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_${s}_fp64_r02b.json 2> gpurun_out/bench_${s}_fp64_r02b.err
  tail -c 1200 gpurun_out/bench_${s}_fp64_r02b.json
done
CMD="python bench.py --precision fp64 --npart-side 256 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_fused_f64 -s 1 -c 1 -o gpurun_out/prof_f64_r02b -f $CMD > gpurun_out/prof_f64_r02b.log 2>&1
ls -la gpurun_out | tail -5
