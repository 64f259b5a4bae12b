#!/bin/bash
# Winograd F(4x4,3x3) path: kernel tests, model parity, cfg2 A/B (BBDM_WINOGRAD=0 vs 1) with per-launch dumps.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_winograd.py -q -x -s 2>&1 | tail -25 > gpurun_out/r2_wino_kernels.log; cat gpurun_out/r2_wino_kernels.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -x -s 2>&1 | grep -v "sampling loop time step" | tail -30 > gpurun_out/r2_wino_model.log; cat gpurun_out/r2_wino_model.log
for v in 0 1; do
  BBDM_WINOGRAD=$v timeout 600 python bench.py --config cfg2 --steps 5 --warmup 3 --no-cpu-baseline --graph --dump-convs gpurun_out/r2_convs_cfg2_wino$v.jsonl > gpurun_out/r2_bench_cfg2_wino$v.json 2>> gpurun_out/r2_bench.err
  python -c "import json; d=json.loads(open('gpurun_out/r2_bench_cfg2_wino$v.json').read().strip().splitlines()[-1]); print('cfg2 wino=$v', round(d['ms_per_step'],2), 'ms  graph', d['config']['graph_replay_ms_per_step'], ' frac', round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_step'],2), d['clocks'], d['config']['activation_pool_gb'])"
done
tail -5 gpurun_out/r2_bench.err
