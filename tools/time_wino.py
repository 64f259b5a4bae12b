#!/usr/bin/env python
"""Time the Winograd transform kernels (and the position GEMM) in isolation for the cfg2 layer shapes.
    python tools/time_wino.py [--once]        # --once: one launch per kernel (for ncu)
    python tools/time_wino.py --output-only   # only the output transform, without / with a same-size residual"""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bbdm_b200 import cabi  # noqa: E402

SHAPES = [(16, 64, 64, 1024, 1024), (16, 128, 128, 512, 512), (16, 256, 256, 512, 512), (16, 128, 128, 1536, 512)]


def timeit(fn, n):
    for _ in range(2 if n > 1 else 0):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    once = "--once" in sys.argv
    be = cabi.CudaBackend()
    dev = "cuda"
    rows = []
    if "--output-only" in sys.argv:
        for B, H, W, Cout in ((16, 128, 128, 512), (16, 64, 64, 1024)):
            th, tw, mt, ok = be.wino_geometry(B, H, W)
            m = torch.randn((36, mt, Cout), device=dev)
            out, res = torch.empty((B, H, W, Cout), device=dev), torch.randn((B, H, W, Cout), device=dev)
            part = torch.empty((B * th, Cout, 2), device=dev)
            t0 = timeit(lambda: be.wino_output(m, B=B, H=H, W=W, Cout=Cout, out=out, stats_partial=part), 10)
            t1 = timeit(lambda: be.wino_output(m, B=B, H=H, W=W, Cout=Cout, out=out, stats_partial=part, residual=res,
                                               res_mode=cabi.RES_SAME), 10)
            gb = (36 * mt * Cout * 4 + B * H * W * Cout * 4) / 1e9
            print(json.dumps({"B": B, "H": H, "W": W, "Cout": Cout, "tiles": mt, "wino_output_ms": t0,
                              "wino_output_tbps": gb / t0, "wino_output_residual_ms": t1,
                              "wino_output_residual_tbps": (gb + B * H * W * Cout * 4 / 1e9) / t1}))
        be.check_fault()
        return
    for B, H, W, Cin, Cout in (SHAPES[1:2] if once else SHAPES):
        x = torch.randn(B, H, W, Cin, device=dev)
        mean, rstd = torch.zeros(B, 32, device=dev), torch.ones(B, 32, device=dev)
        gamma, beta = torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev)
        th, tw, mt, ok = be.wino_geometry(B, H, W)
        vh = torch.empty((36, mt, Cin), dtype=torch.float16, device=dev)
        vl = torch.empty_like(vh)
        uh = torch.randn(36, Cout, Cin, device=dev).to(torch.float16)
        ul = (0.001 * torch.randn(36, Cout, Cin, device=dev)).to(torch.float16)
        m = torch.empty((36, mt, Cout), device=dev)
        out = torch.empty((B, H, W, Cout), device=dev)
        part = torch.empty((B * th, Cout, 2), device=dev)
        n = 1 if once else 10
        t_in = timeit(lambda: be.wino_input(x, None, groups=32, mean=mean, rstd=rstd, gamma=gamma, beta=beta, silu=True,
                                            v_hi=vh, v_lo=vl), n)
        t_g = timeit(lambda: be.conv_umma(B=36, H=mt // 16, W=16, Cin=Cin, Cout=Cout, taps=1, a_hi=vh, a_lo=vl, w_hi=uh,
                                          w_lo=ul, out=m, passes=3, weights_per_image=True, operand_f16=True), n)
        t_o = timeit(lambda: be.wino_output(m, B=B, H=H, W=W, Cout=Cout, out=out, stats_partial=part), n)
        gb_in = (B * H * W * Cin * 4 + 36 * mt * Cin * 4) / 1e9
        gb_out = (36 * mt * Cout * 4 + B * H * W * Cout * 4) / 1e9
        rows.append({"B": B, "H": H, "W": W, "Cin": Cin, "Cout": Cout, "tiles": mt,
                     "wino_input_ms": t_in, "wino_input_tbps": gb_in / t_in, "gemm_ms": t_g,
                     "gemm_algo_tflops": 2.0 * B * H * W * Cout * 9 * Cin / t_g / 1e9,
                     "wino_output_ms": t_o, "wino_output_tbps": gb_out / t_o})
        print(json.dumps(rows[-1]))
        del x, vh, vl, uh, ul, m, out
        torch.cuda.empty_cache()
    be.check_fault()


if __name__ == "__main__":
    main()
