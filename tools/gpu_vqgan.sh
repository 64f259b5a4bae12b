#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vqgan.py -q -m gpu -s 2>&1 | tail -40 > gpurun_out/tests_vqgan.log; cat gpurun_out/tests_vqgan.log
