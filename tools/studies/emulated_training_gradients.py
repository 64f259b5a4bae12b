#!/usr/bin/env python
"""CPU study: one training step of a full-size UNet architecture (default: LBBDM-f4, BASELINE configs[2]) with the
native autograd Functions of bbdm_b200/train.py running on the kernel EMULATION of tests/_emu_backend.py (test
infrastructure: fp16/bf16 operand splitting and the Winograd transforms reproduced in torch), against the stock
PyTorch graph of the same modules.  Prints the per-parameter gradient deviations in module order.

This is the check that located the fp16 range problem of the Winograd data gradient at real loss-gradient magnitudes
(DESIGN section 3): per-Function tests feed unit-scale dY and cannot see it.

    python tools/studies/emulated_training_gradients.py [lbbdm_f4|lbbdm_f8|lbbdm_f16|mid_pixel] [--batch 2] [--min-tiles 16]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from _emu_backend import EmuBackend  # noqa: E402
from _recipe import UNET_CONFIGS, bb_namespace, fill_state_dict, rel_dev  # noqa: E402

import bbdm_b200.unet as U  # noqa: E402
from bbdm_b200 import train  # noqa: E402
from bbdm_b200.bridge import BridgeOps  # noqa: E402
from model.BrownianBridge.BrownianBridgeModel import BrownianBridgeModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default="lbbdm_f4")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--min-tiles", type=int, default=16, help="Winograd tile threshold (512 in the product; lowered "
                    "so that a small batch takes the route the benchmark batch takes)")
    a = ap.parse_args()
    emu = EmuBackend()
    BridgeOps.backend_factory = staticmethod(lambda: emu)
    train.set_backend(emu)
    train.WINO_MIN_TILES = a.min_tiles
    torch.set_num_threads(min(32, os.cpu_count() or 8))
    cfg = UNET_CONFIGS[a.config]
    net = BrownianBridgeModel(bb_namespace(cfg)).train()
    shapes = {k: tuple(v.shape) for k, v in net.denoise_fn.state_dict().items()}
    net.denoise_fn.load_state_dict(fill_state_dict(shapes, seed=1234))
    g = torch.Generator().manual_seed(3)
    c, s = cfg["out_channels"], cfg["image_size"]
    x = torch.randn(a.batch, c, s, s, generator=g).clamp_(-1, 1)
    y = torch.randn(a.batch, c, s, s, generator=g).clamp_(-1, 1)
    nz = torch.randn(a.batch, c, s, s, generator=g)
    t = torch.randint(0, 1000, (a.batch,), generator=g)
    ctx = None if cfg["condition_key"] == "nocond" else y
    res = {}
    for native in (True, False):
        U.NATIVE_TRAIN_CONV = native
        net.zero_grad(set_to_none=True)
        emu.calls.clear()
        t0 = time.time()
        loss, _ = net.p_losses(x, y, ctx, t, nz)
        loss.backward()
        print(f"{'native (emulated kernels)' if native else 'stock graph'}: loss {float(loss):.7f}, {time.time() - t0:.1f} s, "
              f"backend calls: {sorted(set(emu.calls))}")
        res[native] = {n: p.grad.detach().clone() for n, p in net.denoise_fn.named_parameters()}
    devs = [(float(rel_dev(res[True][n], res[False][n])), n) for n in res[False]]
    for d, n in devs:
        print(f"{d:.3e}  {n}  {tuple(res[False][n].shape)}")
    d, n = max(devs)
    print(f"worst: {d:.3e} {n}")


if __name__ == "__main__":
    main()
