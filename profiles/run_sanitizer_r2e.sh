#!/bin/bash
# compute-sanitizer memcheck over the --enemy_comm paths (prey as agent row N in the env step / obs gather, index
# encoders, operand preparation, BPTT) and the e2e host-path changes.  bash profiles/run_sanitizer_r2e.sh (under gpurun)
OUT=gpurun_out/r2e_compute_sanitizer.txt

echo "# compute-sanitizer memcheck, enemy_comm: pytest -m gpu -k enemy (envs, rollout, grad)" > $OUT
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_envs.py tests/test_gpu_rollout.py tests/test_gpu_grad.py -m gpu -k enemy -q --timeout 800 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -6 >> $OUT
cat $OUT
