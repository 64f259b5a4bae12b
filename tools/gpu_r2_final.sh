#!/bin/bash
# Round-2 closing run (1 GPU, ~6 min): full GPU suite, smoke, default bench line, per-launch dump, ncu launch list
# and one --set full capture of conv_umma on the final build (roofline.traffic source).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -v "sampling loop time step" | tail -80 > gpurun_out/r2_tests_full.log; tail -3 gpurun_out/r2_tests_full.log
grep -E "spatial_rescale" gpurun_out/r2_tests_full.log | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_cfg2_default.json 2> gpurun_out/r2_bench_default.err; tail -c 500 gpurun_out/r2_bench_cfg2_default.json; tail -2 gpurun_out/r2_bench_default.err
timeout 600 python bench.py --config cfg2 --steps 10 --warmup 3 --no-cpu-baseline --graph --dump-convs gpurun_out/r2_convs_cfg2_final.jsonl > gpurun_out/r2_bench_cfg2_final.json 2>> gpurun_out/r2_bench_default.err
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_cfg2_final.json').read().strip().splitlines()[-1]); print('cfg2', round(d['ms_per_step'],3), 'ms graph', d['config']['graph_replay_ms_per_step'], 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_cfg2_final.csv python bench.py --config cfg2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch_cfg2.log 2>&1; echo "launch list rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 95 -c 4 -f -o gpurun_out/r2_prof_conv_umma_cfg2 python bench.py --config cfg2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_full_conv.log 2>&1; echo "ncu conv rc $?"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
