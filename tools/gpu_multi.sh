#!/bin/bash
# usage: tools/gpu_multi.sh N  -- NCCL test (N=2 only) + the default bench on N GPUs
N=$1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then ( timeout 900 python -m pytest tests/test_nccl_two_gpus.py -m gpu -x -q 2>&1 | tail -5 ) | tee gpurun_out/pytest_nccl.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_512_${N}gpu.json 2> gpurun_out/bench_512_${N}gpu.err
tail -c 2600 gpurun_out/bench_512_${N}gpu.json
