#!/bin/bash
# Quick check after a kernel change (run under gpurun): parity tests + short bench lines at 256^3 and 512^3.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) | tee gpurun_out/pytest_gpu.log
for s in 256 512; do timeout 300 python bench.py --npart-side $s --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pps %.4g ms %.2f'%(d['value'], d['ms_per_step']), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'frac %.3f'%d['roofline']['frac'])"; done
