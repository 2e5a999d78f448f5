#!/bin/bash
# round 2, call l (2 GPUs): second walk pass through active lists, LET stream with priority: tests + 2-GPU bench line
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_r02l.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("N", d["n_gpus"], "pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],1), "mom %.2e"%d["momentum_residual"])'
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_512_2gpu_r02l.json 2> gpurun_out/bench_512_2gpu_r02l.err; tail -1 gpurun_out/bench_512_2gpu_r02l.json | python -c "$Q"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --npart-side 256 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_256_2gpu_r02l.json 2> gpurun_out/bench_256_2gpu_r02l.err; tail -1 gpurun_out/bench_256_2gpu_r02l.json | python -c "$Q"
