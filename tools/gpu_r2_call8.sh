#!/bin/bash
# training: Winograd fwd/dgrad A/B at the cfg3 UNet micro-step, GPU training tests, cfg4/cfg5 --ends, 1-GPU LBBDM training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_vqgan.py -q -x 2>&1 | tail -4
for w in 0 1; do
  BBDM_WINOGRAD_TRAIN=$w timeout 600 python - <<PY
import json, sys, os
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import torch, bench
from bench_train import run
r = run("native", bench.CONFIGS["cfg3"], steps=4, warmup=2)
print("train cfg3 unet microstep wino=$w", json.dumps(r))
open("gpurun_out/r2_train_cfg3_wino$w.json", "w").write(json.dumps(r))
PY
done
timeout 900 python tools/bench_train_ddp.py --steps 4 --warmup 2 > gpurun_out/r2_train_ddp_1.json 2> gpurun_out/r2_train_ddp_1.err; tail -c 900 gpurun_out/r2_train_ddp_1.json; tail -3 gpurun_out/r2_train_ddp_1.err
for cfg in cfg4 cfg5; do
  timeout 900 python bench.py --config $cfg --ends > gpurun_out/r2_bench_${cfg}_ends.json 2> gpurun_out/r2_bench_${cfg}_ends.err
  python -c "import json; d=json.loads(open('gpurun_out/r2_bench_${cfg}_ends.json').read().strip().splitlines()[-1]); c=d['config']; print('$cfg ends', round(d['value'],2), 'steps/s', {k: round(c[k],1) for k in ('ms_per_batch','encode_ms','decode_ms','loop_ms','images_per_s')})" || tail -5 gpurun_out/r2_bench_${cfg}_ends.err
done
nvidia-smi --query-gpu=memory.used --format=csv
