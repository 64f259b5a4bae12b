#!/bin/bash
mkdir -p gpurun_out
for c in 2 4 8; do
  echo "chunk=$c"
  BBDM_WINO_CHUNK=$c timeout 300 python tools/time_wino.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['Cin'], d['Cout'], d['tiles'], 'gemm', round(d['gemm_ms'],3), round(d['gemm_algo_tflops']))
"
  BBDM_WINO_CHUNK=$c timeout 600 python -m pytest tests/test_gpu_winograd.py tests/test_gpu_model.py -q -x -s -k "chain or fixture" 2>&1 | grep -E "wino chain|cfg2 256|\[cfg1\]|lbbdm_f16\]|passed|failed" | cut -c1-200
  BBDM_WINO_CHUNK=$c timeout 600 python bench.py --config cfg2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg2_chunk$c.json 2>> gpurun_out/r2_bench.err
  python -c "import json; d=json.loads(open('gpurun_out/r2_bench_cfg2_chunk$c.json').read().strip().splitlines()[-1]); print('cfg2 chunk=$c', round(d['ms_per_step'],2), 'ms frac', round(d['roofline']['frac'],4), d['clocks'])"
done
