#!/bin/bash
# round 2, first call: parity tests of the round-1 state + ADVICE fixes, default bench line with the parity block
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_r02a.log
timeout 900 python bench.py > gpurun_out/bench_512_r02a.json 2> gpurun_out/bench_512_r02a.err; tail -c 1500 gpurun_out/bench_512_r02a.json
