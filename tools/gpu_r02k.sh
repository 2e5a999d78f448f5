#!/bin/bash
# round 2, call k (2 GPUs): parity tests incl. the NCCL transport (LET exchange behind the first walk pass, PM all-reduce), 2-GPU bench line
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus_r02k.txt
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu_r02k.log
( timeout 900 python -m pytest tests/test_nccl_two_gpus.py -m gpu -x -q -s 2>&1 | grep -E "NCCL|passed|failed|Error|error|skipped" | tail -20 ) | tee gpurun_out/pytest_nccl_r02k.log
Q='import json,sys; d=json.loads(sys.stdin.read()); print("N", d["n_gpus"], "pps %.4g ms %.2f"%(d["value"], d["ms_per_step"]), {k:round(v,2) for k,v in d["phases_ms"].items()}, "frac %.3f"%d["roofline"]["frac"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],1), "mom %.2e"%d["momentum_residual"])'
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_512_1gpu_r02k.json 2> gpurun_out/bench_512_1gpu_r02k.err; tail -1 gpurun_out/bench_512_1gpu_r02k.json | python -c "$Q"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --rebalance 2 > gpurun_out/bench_512_2gpu_r02k.json 2> gpurun_out/bench_512_2gpu_r02k.err; tail -1 gpurun_out/bench_512_2gpu_r02k.json | python -c "$Q"
tail -3 gpurun_out/bench_512_2gpu_r02k.err
