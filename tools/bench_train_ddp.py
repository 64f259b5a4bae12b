#!/usr/bin/env python
"""End-to-end LBBDM-f4 training step under DDP (BASELINE configs[2]): images [32,3,256,256] per rank ->
two no-grad VQGAN encodes (LatentBrownianBridgeModel.forward, reference :57-62) -> q_sample -> UNet forward ->
L1 loss -> backward (torch DDP: bucketed NCCL allreduce of the 948 MB of UNet gradients over NVLink, exactly what
runners/BaseRunner.py:76 wraps) -> Adam.  One process per GPU:

    python tools/bench_train_ddp.py                                   # 1 GPU (no collective)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_train_ddp.py [--steps 6] [--warmup 3] [--profile]

Rank 0 prints ONE JSON line: ms per micro-step (CUDA events, barrier + synchronize on both sides, MAX over ranks),
samples/s over all ranks, the same step with the allreduce suppressed (DDP.no_sync: what the step would cost with
free communication => exposed communication time), gradient identity across ranks, and -- with --profile -- how
much of the NCCL kernel time ran concurrently with compute kernels (torch.profiler kernel timeline, rank 0).
"""
import argparse
import contextlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

DD = dict(double_z=False, z_channels=3, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 4),
          num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def build(dev, native=True):
    import argparse as ap
    import bbdm_b200.unet as U
    from model.BrownianBridge.LatentBrownianBridgeModel import LatentBrownianBridgeModel
    U.NATIVE_TRAIN_CONV = native
    cfg = bench.CONFIGS["cfg3"]
    ns = bench.namespace(cfg["unet"], cfg["sample_step"])
    ns.VQGAN = ap.Namespace(params=ap.Namespace(ckpt_path=None, embed_dim=3, n_embed=8192, ddconfig=ap.Namespace(**DD),
                                                lossconfig=ap.Namespace(target="torch.nn.Identity")))
    net = LatentBrownianBridgeModel(ns).train()
    bench.init_weights(net.denoise_fn)
    torch.manual_seed(4321)
    with torch.no_grad():
        for n, p in net.vqgan.named_parameters():
            if p.dim() >= 2:
                p.normal_(0, 0.02)
    return net.to(dev), cfg


def overlap_from_profile(prof):
    """(nccl kernel ms, ms of it overlapped by compute kernels on other streams) from the CUDA kernel timeline."""
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    nccl = [(e.time_range.start, e.time_range.end) for e in ev if "nccl" in e.name.lower()]
    comp = sorted((e.time_range.start, e.time_range.end) for e in ev if "nccl" not in e.name.lower() and "memcpy" not in e.name.lower())
    merged = []
    for s, t in comp:
        if merged and s <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], t)
        else:
            merged.append([s, t])
    tot = ov = 0.0
    for s, t in nccl:
        tot += t - s
        for a, b in merged:
            if b <= s:
                continue
            if a >= t:
                break
            ov += min(t, b) - max(s, a)
    return tot / 1e3, ov / 1e3, len(nccl)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--library", action="store_true", help="stock PyTorch TF32 graph instead of the native kernels")
    ap.add_argument("--torch-adam", action="store_true", help="torch.optim.Adam instead of bbdm_b200.optim.FusedAdam")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if a.library:
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
        os.environ["BBDM_NATIVE_VQGAN"] = "0"
    net, cfg = build(dev, native=not a.library)
    if a.library:
        # the frozen autoencoder on stock PyTorch kernels too (the repo's VQModel container has no forward of its own:
        # tools/bench_vqgan.py holds the library-path restatement used for the library rows)
        import types
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import bench_vqgan
        net.encode = types.MethodType(lambda self, x, cond=True, normalize=None: bench_vqgan.lib_encode(self.vqgan, x), net)
    B = cfg["batch"]
    x = bench.synth((B, 3, 256, 256), 100 + rank).to(dev)            # different data per rank, same seed for t/noise (Q6)
    xc = bench.synth((B, 3, 256, 256), 200 + rank).to(dev)
    if a.library or a.torch_adam:
        opt = torch.optim.Adam(net.get_parameters(), lr=1e-4, betas=(0.9, 0.999))
    else:
        from bbdm_b200.optim import FusedAdam                       # one multi-tensor launch per step
        opt = FusedAdam(net.get_parameters(), lr=1e-4, betas=(0.9, 0.999))
    model = net
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], output_device=local)

    def step(sync=True):
        opt.zero_grad(set_to_none=True)
        ctx = contextlib.nullcontext() if (sync or world == 1) else model.no_sync()
        with ctx:
            loss, _ = model(x, xc)
            loss.backward()
        opt.step()
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, sync=True):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loss = step(sync)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / n
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, float(loss)

    torch.manual_seed(1234)
    for _ in range(a.warmup):
        step()
    ms, loss = timed(a.steps)
    row = {"what": "LBBDM-f4 training micro-step: 2 VQGAN encodes (no grad) + q_sample + UNet fwd/bwd + DDP gradient allreduce + Adam",
           "impl": "stock PyTorch TF32 (library)" if a.library else "bbdm_b200 native kernels (split-bf16 x3)",
           "optimizer": type(opt).__name__, "n_gpus": world, "batch_per_gpu": B, "ms_per_micro_step": ms, "micro_steps_per_s": 1e3 / ms,
           "samples_per_s": world * B * 1e3 / ms, "loss": loss, "steps": a.steps, "warmup": a.warmup,
           "grad_bytes_allreduced": sum(p.numel() for p in net.get_parameters()) * 4,
           "max_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    if world > 1:
        ms_nosync, _ = timed(a.steps, sync=False)
        row["ms_per_micro_step_no_allreduce"] = ms_nosync
        row["exposed_communication_ms"] = ms - ms_nosync
        step()                                                      # a synced step, so the gradients below are the averaged ones
        g = torch.cat([p.grad.flatten()[:1000] for p in net.get_parameters()][:20]).double()
        lo, hi = g.clone(), g.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        row["grads_identical_across_ranks"] = bool(torch.equal(lo, hi))
    if a.profile:
        from torch.profiler import ProfilerActivity, profile
        barrier()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        if rank == 0:
            tot, ov, n = overlap_from_profile(prof)
            row["nccl_kernel_ms"], row["nccl_ms_overlapped_with_compute"], row["nccl_kernels"] = tot, ov, n
            row["nccl_overlap_frac"] = (ov / tot) if tot else None
    if rank == 0:
        print(json.dumps(row))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
