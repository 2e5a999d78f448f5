#!/bin/bash
# round 2, call z (N GPUs): one 512^3 line on N GPUs with the judged code (the driver's scaling run repeats all N)
mkdir -p gpurun_out
N=$1; T=${2:-r02z}
Q='import json,sys; d=json.loads(sys.stdin.read()); print("N", d["n_gpus"], d["config"]["npart"], d["dtype"], "pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "step %.3f"%d["roofline"]["whole_step"]["frac"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],1), "ranks", [round(x,2) for x in d["per_rank"]["ms_per_step"]])'
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_512_${N}gpu_$T.json 2> gpurun_out/bench_512_${N}gpu_$T.err; tail -1 gpurun_out/bench_512_${N}gpu_$T.json | python -c "$Q"
