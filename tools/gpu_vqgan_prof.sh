#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_vqgan.csv \
   python tools/bench_vqgan.py 8 --profile > gpurun_out/ncu_vqgan.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_vqgan.log
