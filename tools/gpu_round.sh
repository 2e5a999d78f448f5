#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_r01i.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke_r01i.log
python bench.py > gpurun_out/bench_512_r01i.json 2> gpurun_out/bench_512_r01i.err; tail -c 1200 gpurun_out/bench_512_r01i.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01i.json 2> gpurun_out/bench_ref_r01i.err; tail -c 300 gpurun_out/bench_ref_r01i.json
timeout 900 bash tools/profile.sh r01i 256
