#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -x -k "attention or mid_pixel or cfg1 or lbbdm" 2>&1 | tail -4
timeout 600 python bench.py --config cfg2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print("steps/s", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"])
PY
python tools/time_attention.py | tee gpurun_out/attention_time.json; python tools/time_attention.py 8 256 1024 16
