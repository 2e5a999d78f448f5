#!/bin/bash
# round 2, call i: fully deferred tree build (one intermediate sort + final sort): parity tests, switch-depth sweep at 512^3 and 256^3
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_r02i.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"])'
for t in 1024 64 16384 134217728 8; do echo "== PN2_TREE_TOP_TARGET=$t"; PN2_TREE_TOP_TARGET=$t timeout 300 python bench.py --npart-side 512 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "$Q"; done 2>&1 | tee gpurun_out/sweep_tree_r02i.log
for t in 1024 16777216; do echo "== 256^3 PN2_TREE_TOP_TARGET=$t"; PN2_TREE_TOP_TARGET=$t timeout 300 python bench.py --npart-side 256 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "$Q"; done 2>&1 | tee -a gpurun_out/sweep_tree_r02i.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02i.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/launches_r02i.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_r02i.csv')))
for i,r in enumerate(rows):
    if r and r[0]=="ID": h=r; st=i+1; break
ik,iv=h.index("Kernel Name"),h.index("Metric Value")
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[st:]:
    if len(r)<=iv: continue
    try: v=float(r[iv].replace(",",""))
    except ValueError: continue
    k=r[ik].split("(")[0][:60]; agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:22]: print(f"{k:62s} n={v[0]:5d} total={v[1]/1e6:9.3f} ms share={v[1]/tot*100:5.1f}%")
PY
