#!/bin/bash
# compute-sanitizer over the small parity tests (run under gpurun): memcheck on Mode B / migration / integrator,
# racecheck (shared-memory hazards of the fused walk + P2P kernel) on one small Mode B case.
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_mode_b.py::test_small_tree_lists_forces" tests/test_gpu_mode_b.py::test_empty_and_tiny tests/test_gpu_migrate.py tests/test_integrator.py -m gpu -x -q > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_mode_b.py::test_small_tree_lists_forces[t04]" -m gpu -x -q > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitize_racecheck.log
