#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
bash tools/sweep.sh "-DWS_U=2;-DWS_U=4 -DWS_REGS_WALK=104 -DWS_REGS_P2P=152;-DWS_U=4 -DWS_REGS_WALK=120 -DWS_REGS_P2P=136;-DWS_U=2 -DWS_CTAS_PER_SM=3 -DWS_REGS_P2P=88 -DWS_REGS_WALK=72" 256 2>&1 | tee gpurun_out/sweep_ws.log
