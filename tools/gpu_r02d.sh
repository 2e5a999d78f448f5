#!/bin/bash
# round 2, call d: deferred-top-level tree build: parity tests + bench lines (256^3, 512^3), FP64 occupancy 3
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_r02d.log
for s in 256 512; do timeout 300 python bench.py --npart-side $s --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pps %.4g ms %.2f'%(d['value'], d['ms_per_step']), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'frac %.3f'%d['roofline']['frac'])"; done 2>&1 | tee gpurun_out/bench_quick_r02d.log
for t in 1024 16384 65536; do echo "== PN2_TREE_TOP_TARGET=$t"; PN2_TREE_TOP_TARGET=$t timeout 300 python bench.py --npart-side 512 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pps %.4g ms %.2f'%(d['value'], d['ms_per_step']), {k:round(v,2) for k,v in d['phases_ms'].items()})"; done 2>&1 | tee -a gpurun_out/bench_quick_r02d.log
for v in "-DF64_MIN_BLOCKS=3" "-DF64_MIN_BLOCKS=2"; do
  touch photons-2.0_b200/csrc/pn2_walk.cu
  make -s -C photons-2.0_b200/csrc EXTRA="$v" > /dev/null 2>&1
  echo "== EXTRA=$v"
  timeout 300 python bench.py --precision fp64 --npart-side 256 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pps %.4g ms %.2f'%(d['value'], d['ms_per_step']), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'frac %.3f'%d['roofline']['frac'])"
done 2>&1 | tee gpurun_out/sweep_f64_r02d.log
