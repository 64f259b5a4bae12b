#!/bin/bash
# round 2, call 1: full GPU suite on the default build (incl. the new full-size cfg2 fixture), then A/B of the
# warp-uniform MMA issue build (bbdm_b200/libbbdm_b200_uniform.so, -DBBDM_UNIFORM_ISSUE) on cfg2 / cfg1 / cfg5.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -v "sampling loop time step" | tail -40 > gpurun_out/r2_tests_default.log; tail -5 gpurun_out/r2_tests_default.log
grep -n "cfg2 256" gpurun_out/r2_tests_default.log
U=$PWD/bbdm_b200/libbbdm_b200_uniform.so
BBDM_LIB=$U timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_training.py -q -x 2>&1 | tail -4 > gpurun_out/r2_tests_uniform.log; cat gpurun_out/r2_tests_uniform.log
for cfg in cfg2 cfg1 cfg5; do
  for v in default uniform; do
    if [ $v = uniform ]; then export BBDM_LIB=$U; else unset BBDM_LIB; fi
    timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --graph --dump-convs gpurun_out/r2_convs_${cfg}_$v.jsonl > gpurun_out/r2_bench_${cfg}_$v.json 2>> gpurun_out/r2_bench.err
    python -c "import json; d=json.loads(open('gpurun_out/r2_bench_${cfg}_$v.json').read().strip().splitlines()[-1]); print('$cfg $v', round(d['ms_per_step'],2), 'ms  graph', d['config']['graph_replay_ms_per_step'], ' frac', round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_step'],2), d['clocks'])"
  done
done
unset BBDM_LIB
tail -5 gpurun_out/r2_bench.err
