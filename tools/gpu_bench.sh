#!/bin/bash
# bench + launch list + one ncu --set full capture of the conv kernel (1 GPU).
mkdir -p gpurun_out
CFG=${1:-cfg2}
timeout 900 python bench.py --config $CFG --steps 5 --warmup 3 > gpurun_out/bench_$CFG.json 2> gpurun_out/bench_$CFG.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench_$CFG.json; tail -n 5 gpurun_out/bench_$CFG.err
timeout 900 python bench.py --config $CFG --steps 5 --warmup 3 --precision bf16 --no-cpu-baseline > gpurun_out/bench_${CFG}_bf16.json 2>> gpurun_out/bench_$CFG.err
echo "bench bf16 exit $?"; tail -c 1500 gpurun_out/bench_${CFG}_bf16.json
if [ "${2:-}" = "ncu" ]; then
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$CFG.csv \
     python bench.py --config $CFG --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$CFG.log 2>&1
  echo "ncu launches exit $?"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 60 -c 3 -f -o gpurun_out/prof_conv_$CFG \
     python bench.py --config $CFG --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$CFG.log 2>&1
  echo "ncu full exit $?"
fi
