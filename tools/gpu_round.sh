#!/bin/bash
# One gpurun call: kernel tests (grouped, each under its own timeout so a hang cannot eat the box),
# model tests, smoke.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name timeout cmd...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $to "$@" > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 6 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run k_basic 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "q_sample or p_sample or layout or gather or gn_stats or prep or pack or conv_direct"
run k_attn 300 python -m pytest tests/test_gpu_kernels.py -q -k "attention"
run k_umma 400 python -m pytest tests/test_gpu_kernels.py -q -k "conv_umma"
run model 600 python -m pytest tests/test_gpu_model.py -q -s
run smoke 300 python __graft_entry__.py smoke
