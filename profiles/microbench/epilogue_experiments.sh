#!/bin/bash
# Ablations of the tensor-core LSTM kernel (profiling experiments; the shipped build defines none of the macros):
# rebuilds the library with -DIC3_TC_EXP_* sets and prints whole-step throughput and the policy_step time.
#   SKIP_MMA no tcgen05.mma | SKIP_TMA no operand copies | SKIP_EPI epilogue only releases the accumulator
#   SKIP_MATH no gate math | SKIP_STORE no h'/c'/partial stores
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["kernels"]["policy_step"]["avg_ms"])'
run() { timeout 200 python bench.py --quick 2>/dev/null | tail -1 | python -c "$P"; }
D=-DIC3_TC_EXP_SKIP_
SETS=("" "${D}MMA ${D}TMA" "${D}MMA ${D}EPI" "${D}EPI" "${D}TMA" "${D}MMA ${D}TMA ${D}EPI" "${D}MMA")
for f in "${SETS[@]}"; do
  IC3_NVCC_EXTRA="$f" python -m ic3net_b200.build --force >/dev/null 2>&1
  echo "== flags: '$f'  single"; run
  echo "== flags: '$f'  pair"; IC3_TC_PAIR=1 run
done
python -m ic3net_b200.build --force >/dev/null 2>&1
