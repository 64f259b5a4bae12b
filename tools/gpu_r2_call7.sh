#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_winograd.py tests/test_gpu_kernels.py -q -k "wino or adam or ema or fp16 or denorm or conv_umma" 2>&1 | tail -5
timeout 300 python tools/time_wino.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['Cin'], d['Cout'], d['tiles'], 'gemm', round(d['gemm_ms'],3), round(d['gemm_algo_tflops']), 'in', round(d['wino_input_ms'],3), 'out', round(d['wino_output_ms'],3))
"
timeout 600 python bench.py --config cfg2 --steps 5 --warmup 3 --no-cpu-baseline --graph --dump-convs gpurun_out/r2_convs_cfg2_c7.jsonl > gpurun_out/r2_bench_cfg2_c7.json 2>> gpurun_out/r2_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_cfg2_c7.json').read().strip().splitlines()[-1])
print('cfg2', round(d['ms_per_step'],2), 'ms graph', d['config']['graph_replay_ms_per_step'], 'e2e', round(d['e2e']['ms_per_step'],2), ' frac', round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_step'],2), d['clocks'])
rows=[json.loads(l) for l in open('gpurun_out/r2_convs_cfg2_c7.jsonl')]
for name in ('wino_input','wino_output'):
    print('  ', name, round(sum(r['ms'] for r in rows if r.get('transform')==name),2), 'ms')
print('   wino gemm', round(sum(r['ms'] for r in rows if r.get('wino')),2), 'ms; direct', round(sum(r['ms'] for r in rows if 'transform' not in r and not r.get('wino')),2))
PY
for cfg in cfg3 cfg4; do
  timeout 900 python bench.py --config $cfg --ends > gpurun_out/r2_bench_${cfg}_ends.json 2> gpurun_out/r2_bench_${cfg}_ends.err
  python -c "import json; d=json.loads(open('gpurun_out/r2_bench_${cfg}_ends.json').read().strip().splitlines()[-1]); c=d['config']; print('$cfg ends', round(d['value'],2), 'steps/s', {k: round(c[k],1) for k in ('ms_per_batch','encode_ms','decode_ms','loop_ms','images_per_s')})" || tail -5 gpurun_out/r2_bench_${cfg}_ends.err
done
tail -3 gpurun_out/r2_bench.err
