#!/bin/bash
# compute-sanitizer over the small parity tests (run under gpurun): memcheck on Mode B (all three arithmetic modes, both tree
# builders, ragged kernel), multi-rank (two walk passes, active lists), PM, migration, integrator, snapshot; racecheck
# (shared-memory hazards) on the fused walk + P2P kernels and the tree builder's block table.   usage: tools/gpu_sanitize.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_mode_b.py::test_small_tree_lists_forces" tests/test_gpu_mode_b.py::test_empty_and_tiny tests/test_gpu_mode_b.py::test_clustered_and_ragged "tests/test_gpu_multirank.py::test_small_vs_oracle" tests/test_gpu_pm.py tests/test_gpu_migrate.py tests/test_integrator.py tests/test_snapshot.py -m gpu -x -q > gpurun_out/sanitize_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_memcheck_$TAG.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_mode_b.py::test_small_tree_lists_forces[t04]" "tests/test_gpu_multirank.py::test_small_vs_oracle[t04-2]" -m gpu -x -q > gpurun_out/sanitize_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitize_racecheck_$TAG.log
