#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_512_r01h.json 2> gpurun_out/bench_512_r01h.err; tail -c 2500 gpurun_out/bench_512_r01h.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01h.json 2> gpurun_out/bench_ref_r01h.err; tail -c 600 gpurun_out/bench_ref_r01h.json
timeout 900 bash tools/profile.sh r01h 256
