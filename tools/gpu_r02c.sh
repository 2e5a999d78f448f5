#!/bin/bash
# round 2, call c: PM + snapshot device tests, FP64 kernel occupancy sweep
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_r02c.log
( timeout 600 python -m pytest tests/test_gpu_pm.py tests/test_snapshot.py -m gpu -q -s 2>&1 | grep -E "rms rel|density|pair|passed|failed|Error|error" | tail -20 ) | tee gpurun_out/pytest_gpu_r02c_pm.log
for v in "-DF64_MIN_BLOCKS=4" "-DF64_MIN_BLOCKS=5" "-DF64_MIN_BLOCKS=6" "-DF64_MIN_BLOCKS=5 -DF64_NST=2" "-DF64_MIN_BLOCKS=5 -DF64_NST=8"; do
  touch photons-2.0_b200/csrc/pn2_walk.cu
  make -s -C photons-2.0_b200/csrc EXTRA="$v" > /dev/null 2>&1
  echo "== EXTRA=$v"
  timeout 300 python bench.py --precision fp64 --npart-side 256 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pps %.4g ms %.2f'%(d['value'], d['ms_per_step']), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'frac %.3f'%d['roofline']['frac'])"
done 2>&1 | tee gpurun_out/sweep_f64_r02c.log
