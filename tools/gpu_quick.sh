#!/bin/bash
# tests (all) + bench split3/bf16 + launch list (no full ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/tests.log; cat gpurun_out/tests.log
CFG=${1:-cfg2}
timeout 900 python bench.py --config $CFG --steps 5 --warmup 3 > gpurun_out/bench_$CFG.json 2> gpurun_out/bench_$CFG.err
echo "bench exit $?"; python - <<PY
import json
for f in ("gpurun_out/bench_$CFG.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "steps/s", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], "conv_ms", d["roofline"]["kernel_ms_per_step"], "launches", d["gpu_launches"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
tail -n 3 gpurun_out/bench_$CFG.err
timeout 900 python bench.py --config $CFG --steps 5 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_${CFG}_bf16.json 2>> gpurun_out/bench_$CFG.err
echo "bench bf16 exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_${CFG}_bf16.json').read().strip().splitlines()[-1]); print('bf16 steps/s', d['value'], 'ms', d['ms_per_step'], 'roof', d['roofline']['frac'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_$CFG.csv \
     python bench.py --config $CFG --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$CFG.log 2>&1
echo "ncu launches exit $?"
