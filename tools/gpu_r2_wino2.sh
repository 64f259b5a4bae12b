#!/bin/bash
# Winograd path, second pass: all kernel tests, input-transform variants (1 or 2 channels per thread), small configs.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_winograd.py -q -s 2>&1 | tail -25 > gpurun_out/r2_wino_kernels.log; cat gpurun_out/r2_wino_kernels.log
for vec in 1 2; do
  BBDM_WINO_IN_VEC=$vec timeout 600 python bench.py --config cfg2 --steps 5 --warmup 3 --no-cpu-baseline --dump-convs gpurun_out/r2_convs_cfg2_winovec$vec.jsonl > gpurun_out/r2_bench_cfg2_winovec$vec.json 2>> gpurun_out/r2_bench.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_cfg2_winovec$vec.json').read().strip().splitlines()[-1])
print('cfg2 vec=$vec', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['ms_per_step'],2), ' frac', round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_step'],2), d['clocks'])
rows=[json.loads(l) for l in open('gpurun_out/r2_convs_cfg2_winovec$vec.jsonl')]
for name in ('wino_input','wino_output'):
    print('  ', name, round(sum(r['ms'] for r in rows if r.get('transform')==name),2), 'ms')
print('   wino gemm', round(sum(r['ms'] for r in rows if r.get('wino')),2), 'ms; direct', round(sum(r['ms'] for r in rows if 'transform' not in r and not r.get('wino')),2))
PY
done
for cfg in cfg1 cfg3 cfg5; do
  for w in 0 1; do
    BBDM_WINOGRAD=$w timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --graph > gpurun_out/r2_bench_${cfg}_wino$w.json 2>> gpurun_out/r2_bench.err
    python -c "import json; d=json.loads(open('gpurun_out/r2_bench_${cfg}_wino$w.json').read().strip().splitlines()[-1]); print('$cfg wino=$w', round(d['ms_per_step'],3), 'ms  graph', d['config']['graph_replay_ms_per_step'])"
  done
done
tail -5 gpurun_out/r2_bench.err
