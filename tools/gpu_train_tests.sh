#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -q -m gpu -x 2>&1 | tail -25 > gpurun_out/tests_train.log; cat gpurun_out/tests_train.log
timeout 600 python tools/bench_train.py cfg3 > gpurun_out/train_cfg3.json 2> gpurun_out/train_cfg3.err; tail -3 gpurun_out/train_cfg3.err; cat gpurun_out/train_cfg3.json | tail -3
