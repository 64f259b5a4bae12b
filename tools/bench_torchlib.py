#!/usr/bin/env python
"""GPU *library* baseline (SURVEY section 8d "the real bar"): the same UNet architecture executed by
stock PyTorch kernels (cuDNN conv, ATen GroupNorm/SiLU/softmax, cuBLAS bmm) -- i.e. what the
reference's code path dispatches to on a B200 under torch 2.11 -- at cfg2, beside our numbers.
Uses bbdm_b200.unet.UNetModel._forward_autograd (plain torch ops over identical parameters; it is
checked against the reference by the test-suite) under no_grad, plus the torch-op bridge update.
Rows: fp32 (TF32 off), fp32 with TF32 convs/matmuls (torch 1.12's conv default), bf16 autocast
(+ channels_last)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bbdm_b200.unet as U  # noqa: E402
from bbdm_b200.unet import UNetModel  # noqa: E402

U.NATIVE_TRAIN_CONV = False        # the module forwards on stock PyTorch kernels only (no bbdm_b200 autograd Functions)


def run(mode, cfg, steps=3, warmup=2):
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    torch.backends.cudnn.benchmark = True
    net = UNetModel(**cfg["unet"]).eval()
    bench.init_weights(net)
    net = net.to(dev)
    B, C, S = cfg["batch"], cfg["channels"], cfg["size"]
    x = bench.synth((B, C, S, S), 1).to(dev)
    y = bench.synth((B, C, S, S), 2).to(dev)
    if mode == "bf16":
        net = net.to(memory_format=torch.channels_last)
    t = torch.full((B,), 500, device=dev, dtype=torch.long)

    def step(x):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16"):
            eps = net._forward_autograd(x, t, y).float()
        x0 = x - eps                                   # objective 'grad' bridge update, torch ops
        noise = torch.randn_like(x)
        return 0.5 * x0 + 0.5 * y + 0.1 * (x - 0.5 * x0 - 0.5 * y) + 0.05 * noise

    for _ in range(warmup):
        x = step(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        x = step(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"mode": mode, "ms_per_step": ms, "steps_per_s": 1e3 / ms,
            "unet_tflops_per_s": cfg["flops_per_step"] / ms / 1e9, "max_mem_gb": torch.cuda.max_memory_allocated() / 1e9}


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    cfg = bench.CONFIGS[name]
    out = {"config": cfg["name"], "what": "stock PyTorch 2.11 library path (cuDNN/cuBLAS/ATen), same architecture and weights init", "rows": []}
    for mode in ("fp32", "tf32", "bf16"):
        try:
            out["rows"].append(run(mode, cfg))
        except Exception as e:  # noqa: BLE001
            out["rows"].append({"mode": mode, "error": repr(e)[:300]})
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    print(json.dumps(out))
