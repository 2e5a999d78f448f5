#!/bin/bash
# round 2, call w (8 GPUs, charged eightfold: two bench lines only): 512^3 on 8 GPUs in FP32 and FP64 mode with the judged code (per-rank timings in the line)
mkdir -p gpurun_out
T=${1:-r02w}
Q='import json,sys; d=json.loads(sys.stdin.read()); print("N", d["n_gpus"], d["config"]["npart"], d["dtype"], "pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "step %.3f / %.3f"%(d["roofline"]["whole_step"]["frac"], d["roofline"]["whole_step"]["frac_of_nominal_peak"]), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],1), "mom %.2e"%d["momentum_residual"], "ranks", [round(x,2) for x in d["per_rank"]["ms_per_step"]])'
tr() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@"; }
tr 8 --steps 10 --warmup 3 > gpurun_out/bench_512_8gpu_$T.json 2> gpurun_out/bench_512_8gpu_$T.err; tail -1 gpurun_out/bench_512_8gpu_$T.json | python -c "$Q"
tr 8 --precision fp64 --steps 3 --warmup 2 --no-e2e > gpurun_out/bench_512_8gpu_fp64_$T.json 2> gpurun_out/bench_512_8gpu_fp64_$T.err; tail -1 gpurun_out/bench_512_8gpu_fp64_$T.json | python -c "$Q"
tail -2 gpurun_out/bench_512_8gpu_$T.err
