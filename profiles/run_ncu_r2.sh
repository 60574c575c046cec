#!/bin/bash
# Round-2 ncu captures (run under gpurun, ONE GPU):  bash profiles/run_ncu_r2.sh
# launch lists (gpu__time_duration.sum, cold / serialised: compare SHARES, not absolutes) + `--set full` captures of the
# rollout kernels (index and dense step) and of the BPTT kernels; summarised by profiles/summarize_ncu.py.
set -x
OUT=gpurun_out
NCU="ncu --clock-control none"
# ---- launch lists ----
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/r2_launches_index.csv python bench.py --quick --skip kernels --steps 4 --warmup 3 --obs_mode index --no_graph > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/r2_launches_dense.csv python bench.py --quick --skip kernels --steps 4 --warmup 3 --no_graph > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/r2_launches_bptt.csv python profiles/tools/bptt_profile.py pp_hard_ic3net 8 > /dev/null 2>&1
# ---- full captures ----
$NCU --set full --import-source on -k "regex:lstm_tc_kernel|prep_kernel|heads_finish_kernel|pp_step_kernel" -s 24 -c 4 -f -o $OUT/r2_ncu_index python bench.py --quick --skip kernels --steps 6 --warmup 3 --obs_mode index --no_graph > /dev/null 2>&1
$NCU --set full --import-source on -k "regex:pp_step_kernel|encoder_dense_kernel" -s 18 -c 3 -f -o $OUT/r2_ncu_dense python bench.py --quick --skip kernels --steps 6 --warmup 3 --no_graph > /dev/null 2>&1
$NCU --set full --import-source on -k "regex:bptt_gates_kernel|bptt_dgrad_kernel|bptt_wgrad_kernel|bptt_heads_kernel|bptt_comm_kernel|prep_kernel" -s 40 -c 7 -f -o $OUT/r2_ncu_bptt python profiles/tools/bptt_profile.py pp_hard_ic3net 8 > /dev/null 2>&1
ls -la $OUT/*.ncu-rep
