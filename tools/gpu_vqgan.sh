#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vqgan.py -q -m gpu -s 2>&1 | grep -v "^E  " | tail -40 > gpurun_out/tests_vqgan.log; cat gpurun_out/tests_vqgan.log
timeout 600 python tools/bench_vqgan.py 32 > gpurun_out/vqgan_f4.json 2> gpurun_out/vqgan_f4.err; tail -3 gpurun_out/vqgan_f4.err; cat gpurun_out/vqgan_f4.json
