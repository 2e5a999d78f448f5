#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
bash tools/sweep.sh "-DLEAF_MIN_BLOCKS=5;-DLEAF_MIN_BLOCKS=6;-DLEAF_MIN_BLOCKS=4;-DLEAF_MIN_BLOCKS=7" 256 2>&1 | tee gpurun_out/sweep_fused2.log
