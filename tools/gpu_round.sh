#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_migrate.py tests/test_nccl_two_gpus.py -m gpu -x -q 2>&1 | tail -25 ) | tee gpurun_out/pytest_migrate.log
