#!/bin/bash
mkdir -p gpurun_out
for c in cfg1 cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --graph --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  echo "$c exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_$c.json').read().strip().splitlines()[-1]); print('$c', 'steps/s', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'graph ms', d['config']['graph_replay_ms_per_step'], 'roof', round(d['roofline']['frac'],3), 'conv share', round(d['roofline']['share_of_step'],2))" 2>&1 | tail -1
  tail -n 2 gpurun_out/bench_$c.err
done
