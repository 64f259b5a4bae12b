#!/bin/bash
mkdir -p gpurun_out
for v in 1 2; do echo "vec=$v"; BBDM_WINO_IN_VEC=$v timeout 300 python tools/time_wino.py 2>&1 | tee gpurun_out/r2_time_wino_vec$v.jsonl; done
BBDM_WINO_IN_VEC=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:wino_input -c 1 -f -o gpurun_out/r2_prof_wino_input python tools/time_wino.py --once > gpurun_out/r2_ncu_wino_input.log 2>&1; echo "ncu rc $?"
BBDM_WINO_IN_VEC=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -c 1 -f -o gpurun_out/r2_prof_wino_gemm python tools/time_wino.py --once > gpurun_out/r2_ncu_wino_gemm.log 2>&1; echo "ncu rc $?"
ls -la gpurun_out/*.ncu-rep
