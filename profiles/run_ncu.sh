#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md applied to bench.py (run under gpurun, 1 GPU).
# Usage: bash profiles/run_ncu.sh <round-tag> [extra bench flags]
TAG=${1:-r1}
shift
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --quick --steps 12 --warmup 3 $@"
BI="python bench.py --quick --steps 12 --warmup 3 --obs_mode index $@"
# every launch with its device time (cold-cache, serialised: compare SHARES), dense and index-form rollouts
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches_$TAG.csv $B > $OUT/ncu_launches_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches_${TAG}_index.csv $BI > $OUT/ncu_launches_${TAG}_index.log 2>&1
# full captures in steady state (skip the first launches of each kernel)
ncu --set full --clock-control none --import-source on -k regex:pp_step_kernel -s 8 -c 2 -o $OUT/prof_ppstep_$TAG $B > $OUT/ncu_ppstep_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encoder_dense_kernel -s 4 -c 1 -o $OUT/prof_encoder_$TAG $B > $OUT/ncu_encoder_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:lstm_tc|policy_step_kernel" -s 4 -c 1 -o $OUT/prof_policy_$TAG $B > $OUT/ncu_policy_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:prep_kernel|heads_finish_kernel|heads_kernel" -s 8 -c 2 -o $OUT/prof_prephead_$TAG $B > $OUT/ncu_prephead_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:prep_kernel" -s 6 -c 1 -o $OUT/prof_prep_index_$TAG $BI > $OUT/ncu_prep_index_$TAG.log 2>&1
ls -la $OUT | tail -12
