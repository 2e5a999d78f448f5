#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
./tools/ubench/ubench_p2p 2>&1 | tail -7
bash tools/sweep.sh "-DFUSED_NST=8;-DFUSED_NST=4;-DFUSED_NST=16 -DLEAF_MIN_BLOCKS=4;-DWALK_WARPS=2 -DLEAF_MIN_BLOCKS=10;-DWALK_WARPS=8 -DLEAF_MIN_BLOCKS=2;-DFUSED_NST=4 -DLEAF_MIN_BLOCKS=6" 256 2>&1 | tee gpurun_out/sweep_nst.log
