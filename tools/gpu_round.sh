#!/bin/bash
# One round on a single-GPU box (run under gpurun): parity tests, smoke, the default bench line, the reference arm,
# and the profiles (launch list of the default workload + full ncu captures).  usage: tools/gpu_round.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke_$TAG.log
python bench.py > gpurun_out/bench_512_$TAG.json 2> gpurun_out/bench_512_$TAG.err; tail -c 600 gpurun_out/bench_512_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 200 gpurun_out/bench_ref_$TAG.json
[ "$2" = "noprofile" ] || timeout 900 bash tools/profile.sh $TAG 256
