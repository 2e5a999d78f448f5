#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/ubench_p2p 2>&1 | tail -8 | tee gpurun_out/ubench_p2p.txt
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
bash tools/sweep.sh "-DLEAF_MIN_BLOCKS=7;-DLEAF_MIN_BLOCKS=6;-DLEAF_MIN_BLOCKS=5;-DLEAF_MIN_BLOCKS=5 -DPN2_RDEG=6" 256 2>&1 | tee gpurun_out/sweep_tiles.log
