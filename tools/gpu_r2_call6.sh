#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
for bk in 64 32; do
  echo "BK=$bk"; BBDM_CONV_BK=$bk timeout 300 python tools/time_wino.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['Cin'], d['Cout'], d['tiles'], 'gemm', round(d['gemm_ms'],3), round(d['gemm_algo_tflops']), 'in', round(d['wino_input_ms'],3), 'out', round(d['wino_output_ms'],3))
"
  BBDM_CONV_BK=$bk timeout 600 python bench.py --config cfg2 --steps 5 --warmup 3 --no-cpu-baseline --graph --dump-convs gpurun_out/r2_convs_cfg2_bk$bk.jsonl > gpurun_out/r2_bench_cfg2_bk$bk.json 2>> gpurun_out/r2_bench.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_cfg2_bk$bk.json').read().strip().splitlines()[-1])
print('cfg2 BK=$bk', round(d['ms_per_step'],2), 'ms graph', d['config']['graph_replay_ms_per_step'], 'e2e', round(d['e2e']['ms_per_step'],2), ' frac', round(d['roofline']['frac'],4), 'conv ms', round(d['roofline']['kernel_ms_per_step'],2), d['clocks'])
rows=[json.loads(l) for l in open('gpurun_out/r2_convs_cfg2_bk$bk.jsonl')]
for name in ('wino_input','wino_output'):
    print('  ', name, round(sum(r['ms'] for r in rows if r.get('transform')==name),2), 'ms')
print('   wino gemm', round(sum(r['ms'] for r in rows if r.get('wino')),2), 'ms; direct', round(sum(r['ms'] for r in rows if 'transform' not in r and not r.get('wino')),2))
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_cfg2_v1.csv python bench.py --config cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch_cfg2.log 2>&1; echo "ncu launches rc $?"
tail -3 gpurun_out/r2_bench.err
