#!/bin/bash
# attention kernel: tests + model parity + one ncu --set full capture
mkdir -p gpurun_out

timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -f -o gpurun_out/prof_attention_tc2_cfg2 \
   python bench.py --config cfg2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_attn2.log 2>&1
echo "ncu exit $?"
