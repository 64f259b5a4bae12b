#!/bin/bash
# DDP training evidence: N = $1 GPUs (plus the 1-GPU line when N == 2)
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_vqgan.py -q -x 2>&1 | tail -3
  timeout 600 python tools/bench_train_ddp.py --steps 5 --warmup 3 > gpurun_out/r2_train_ddp_1.json 2> gpurun_out/r2_train_ddp_1.err; tail -c 700 gpurun_out/r2_train_ddp_1.json
  timeout 600 python tools/bench_train_ddp.py --steps 5 --warmup 3 --library > gpurun_out/r2_train_ddp_1_library_tf32.json 2>> gpurun_out/r2_train_ddp_1.err; tail -c 500 gpurun_out/r2_train_ddp_1_library_tf32.json
  timeout 600 python tools/bench_vqgan.py 32 > gpurun_out/r2_vqgan_f4_ends.json 2> gpurun_out/r2_vqgan.err; tail -c 1200 gpurun_out/r2_vqgan_f4_ends.json
fi
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tools/bench_train_ddp.py --steps 5 --warmup 3 --profile > gpurun_out/r2_train_ddp_$N.json 2> gpurun_out/r2_train_ddp_$N.err
echo "ddp $N rc $?"; tail -c 1200 gpurun_out/r2_train_ddp_$N.json; tail -3 gpurun_out/r2_train_ddp_$N.err
