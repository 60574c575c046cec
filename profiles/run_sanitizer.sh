#!/bin/bash
# compute-sanitizer over a subset of the GPU parity tests (tc path incl. tcgen05 kernels, fused table encoder, optimizer).
# Usage (under gpurun): bash profiles/run_sanitizer.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out/compute_sanitizer_$TAG.txt
T="tests/test_gpu_policy.py::test_forward_matches_reference_golden tests/test_gpu_rollout.py::test_dense_and_index_rollouts_are_identical tests/test_gpu_optim.py::test_flat_rmsprop_matches_torch_fixture tests/test_gpu_optim.py::test_rmsprop_argument_checks tests/test_gpu_edges.py"
R="tests/test_gpu_rollout.py::test_dense_and_index_rollouts_are_identical tests/test_gpu_policy.py::test_forward_matches_reference_golden"
{
  echo "# compute-sanitizer, tag $TAG: $T"
  echo "## memcheck"
  timeout 900 compute-sanitizer --tool memcheck python -m pytest $T -x -q -m gpu 2>&1 | grep -v "^$" | tail -8
  echo "## racecheck"
  timeout 900 compute-sanitizer --tool racecheck python -m pytest $R -x -q -m gpu -k "tc or identical" 2>&1 | grep -v "^$" | tail -8
} > $OUT
cat $OUT
