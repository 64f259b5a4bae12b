#!/bin/bash
# Round-2 evidence run (1 GPU): full GPU suite, smoke, bench lines of every config, launch list, ncu --set full
# captures of the conv / Winograd kernels, library baseline.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
[ -n "$SKIP_TESTS" ] || timeout 1200 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -v "sampling loop time step" | tail -60 > gpurun_out/r2_tests_full.log; tail -4 gpurun_out/r2_tests_full.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_cfg2_default.json 2> gpurun_out/r2_bench_default.err; tail -c 600 gpurun_out/r2_bench_cfg2_default.json; tail -2 gpurun_out/r2_bench_default.err
timeout 600 python bench.py --impl reference --steps 1 > gpurun_out/r2_bench_cfg2_reference_arm.json 2>> gpurun_out/r2_bench_default.err; tail -c 400 gpurun_out/r2_bench_cfg2_reference_arm.json
for cfg in cfg2 cfg1 cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --graph --dump-convs gpurun_out/r2_convs_${cfg}_final.jsonl > gpurun_out/r2_bench_${cfg}_final.json 2>> gpurun_out/r2_bench_default.err
  python -c "import json; d=json.loads(open('gpurun_out/r2_bench_${cfg}_final.json').read().strip().splitlines()[-1]); print('$cfg', round(d['ms_per_step'],3), 'ms graph', d['config']['graph_replay_ms_per_step'], 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), d['clocks'])"
done
timeout 600 python bench.py --config cfg2 --steps 5 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bench_cfg2_bf16.json 2>> gpurun_out/r2_bench_default.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_cfg2_final.csv python bench.py --config cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch_cfg2.log 2>&1; echo "launch list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 95 -c 4 -f -o gpurun_out/r2_prof_conv_umma_cfg2 python bench.py --config cfg2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_full_conv.log 2>&1; echo "ncu conv rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wino_ -s 44 -c 2 -f -o gpurun_out/r2_prof_wino_cfg2 python bench.py --config cfg2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_full_wino.log 2>&1; echo "ncu wino rc $?"
timeout 900 python tools/bench_torchlib.py cfg2 > gpurun_out/r2_torchlib_cfg2.json 2> gpurun_out/r2_torchlib.err; tail -c 700 gpurun_out/r2_torchlib_cfg2.json
timeout 600 python tools/bench_train_ddp.py --steps 4 --warmup 2 --library > gpurun_out/r2_train_ddp_1_library_tf32.json 2> gpurun_out/r2_train_lib.err; tail -c 500 gpurun_out/r2_train_ddp_1_library_tf32.json; tail -2 gpurun_out/r2_train_lib.err
ls -la gpurun_out/*.ncu-rep
