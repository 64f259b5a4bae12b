#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_winograd.py tests/test_gpu_kernels.py -q -k "wino or adam or ema or fp16" 2>&1 | tail -8
echo "smem (default)"; timeout 300 python tools/time_wino.py 2>&1 | tee gpurun_out/r2_time_wino_smem.jsonl
timeout 600 python bench.py --config cfg2 --steps 5 --warmup 3 --no-cpu-baseline --graph --dump-convs gpurun_out/r2_convs_cfg2_smem.jsonl > gpurun_out/r2_bench_cfg2_smem.json 2>> gpurun_out/r2_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_cfg2_smem.json').read().strip().splitlines()[-1])
print('cfg2', round(d['ms_per_step'],2), 'ms graph', d['config']['graph_replay_ms_per_step'], 'e2e', round(d['e2e']['ms_per_step'],2), ' frac', round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_step'],2), d['clocks'])
rows=[json.loads(l) for l in open('gpurun_out/r2_convs_cfg2_smem.jsonl')]
for name in ('wino_input','wino_output'):
    print('  ', name, round(sum(r['ms'] for r in rows if r.get('transform')==name),2), 'ms')
print('   wino gemm', round(sum(r['ms'] for r in rows if r.get('wino')),2), 'ms; direct', round(sum(r['ms'] for r in rows if 'transform' not in r and not r.get('wino')),2))
PY
timeout 900 python -m pytest tests/test_gpu_model.py -q -x -s 2>&1 | grep -E "rel dev|passed|failed" | grep -v "sampling loop" | tail -16
tail -3 gpurun_out/r2_bench.err
