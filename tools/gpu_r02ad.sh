#!/bin/bash
# round 2, call ad (1 GPU): the full GPU test suite on the final code
mkdir -p gpurun_out
( timeout -s ABRT 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) | tee gpurun_out/pytest_gpu_r02ad.log
