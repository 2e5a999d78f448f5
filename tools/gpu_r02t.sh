#!/bin/bash
# round 2, call t (8 GPUs; every minute is charged eightfold, so only what needs eight): NCCL parity at 2 / 4 / 8 ranks; 512^3 and 1024^3 on 8 GPUs
# (leaf walk v2, R degree 7, table M2L, bounded grids in the pass over the received trees)
mkdir -p gpurun_out
T=${1:-r02t}
nvidia-smi -L | wc -l | tee gpurun_out/gpus_$T.txt; nproc | tee -a gpurun_out/gpus_$T.txt
( timeout 600 python -m pytest tests/test_nccl_two_gpus.py -m gpu -x -q -s 2>&1 | grep -E "NCCL|passed|failed|Error|error|skipped" | tail -30 ) | tee gpurun_out/pytest_nccl_$T.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("N", d["n_gpus"], d["config"]["npart"], d["dtype"], "pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],1), "mom %.2e"%d["momentum_residual"], "pm", d.get("pm_long_range") and round(d["pm_long_range"]["ms"],2))'
tr() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@"; }
tr 8 --steps 10 --warmup 3 > gpurun_out/bench_512_8gpu_$T.json 2> gpurun_out/bench_512_8gpu_$T.err; tail -1 gpurun_out/bench_512_8gpu_$T.json | python -c "$Q"
tr 8 --npart-side 1024 --steps 3 --warmup 2 > gpurun_out/bench_1024_8gpu_$T.json 2> gpurun_out/bench_1024_8gpu_$T.err; tail -1 gpurun_out/bench_1024_8gpu_$T.json | python -c "$Q"
tail -2 gpurun_out/bench_1024_8gpu_$T.err
