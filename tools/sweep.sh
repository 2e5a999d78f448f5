#!/bin/bash
# usage: tools/sweep.sh "<make EXTRA flags variants separated by ;>" <npart-side>
IFS=';' read -ra V <<< "$1"
SIDE=${2:-256}
for v in "${V[@]}"; do
  touch photons-2.0_b200/csrc/pn2_walk.cu
  make -s -C photons-2.0_b200/csrc EXTRA="$v" > /dev/null 2>&1
  echo "== EXTRA=$v"
  timeout 120 python bench.py --npart-side $SIDE --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pps %.4g ms %.2f'%(d['value'], d['ms_per_step']), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'frac %.3f'%d['roofline']['frac'])"
done
