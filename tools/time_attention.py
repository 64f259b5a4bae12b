#!/usr/bin/env python
"""Time bbdm_attention_tc alone at the cfg2 shape (B=16, T=4096, C=1024, 16 heads) with CUDA events."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bbdm_b200 import cabi
be = cabi.CudaBackend()
B, T, C, heads = (int(a) for a in (sys.argv[1:5] if len(sys.argv) > 4 else (16, 4096, 1024, 16)))
q = torch.randn(B, T, 3 * C, device="cuda")
hi = q.to(torch.bfloat16); lo = (q - hi.float()).to(torch.bfloat16)
o_hi = torch.empty(B, T, C, dtype=torch.bfloat16, device="cuda"); o_lo = torch.empty_like(o_hi)
for _ in range(3):
    be.attention_tc(hi, lo, heads, 0, None, o_hi, o_lo)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    be.attention_tc(hi, lo, heads, 0, None, o_hi, o_lo)
e1.record(); torch.cuda.synchronize()
be.check_fault()
ms = e0.elapsed_time(e1) / 10
fl = 4.0 * B * heads * T * T * (C // heads)
print(json.dumps({"kernel": "attention_tc", "B": B, "T": T, "C": C, "heads": heads, "ms": ms, "algo_tflops": fl / ms / 1e9}))
