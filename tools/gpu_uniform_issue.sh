#!/bin/bash
# Round-2 first step: validate the opt-in warp-uniform MMA issue variant of conv_umma / conv_wgrad.
# Build it HERE first (BBDM_NVCC_DEFINES=-DBBDM_UNIFORM_ISSUE python -m bbdm_b200.build --force), then
#   gpurun --timeout 1200 -- 'bash tools/gpu_uniform_issue.sh'
# runs the whole GPU suite and the cfg2 / cfg1 / cfg5 bench lines in both precision modes for comparison with
# profiles/r01_bench_*_v5/v6.json.  If green and faster: define the macro in bbdm_b200/build.py FLAGS.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/tests_uniform.log; cat gpurun_out/tests_uniform.log
for cfg in cfg2 cfg1 cfg5; do
  timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_uniform_$cfg.json 2>> gpurun_out/bench_uniform.err
  python -c "import json; d=json.loads(open('gpurun_out/bench_uniform_$cfg.json').read().strip().splitlines()[-1]); print('$cfg split3', d['ms_per_step'], 'ms', d['roofline']['frac'])"
done
timeout 600 python bench.py --config cfg2 --steps 5 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_uniform_cfg2_bf16.json 2>> gpurun_out/bench_uniform.err
python -c "import json; d=json.loads(open('gpurun_out/bench_uniform_cfg2_bf16.json').read().strip().splitlines()[-1]); print('cfg2 bf16', d['ms_per_step'], 'ms', d['roofline']['frac'])"
timeout 600 python tools/bench_train.py cfg3 > gpurun_out/train_uniform_cfg3.json 2>> gpurun_out/bench_uniform.err; tail -c 400 gpurun_out/train_uniform_cfg3.json
