#!/bin/bash
# one ncu --set full capture per kernel family (1 GPU), cfg2
mkdir -p gpurun_out
for k in conv_umma attention_tc prep_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 2 -f -o gpurun_out/prof_${k}_cfg2 \
     python bench.py --config cfg2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1
  echo "$k exit $?"
done
