#!/usr/bin/env python
"""HBM roofline of the Brownian-bridge elementwise kernels (q_sample, p_sample) and the layout edge
kernels at the cfg2 tensor size x a large replication factor (so the working set exceeds L2), CUDA
events on the launching stream.  Algorithmic bytes: q_sample 20 B/elem (3 reads + 2 writes),
p_sample 24 B/elem (4 reads + 2 writes)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bbdm_b200 import cabi  # noqa: E402
from bbdm_b200.schedule import bridge_buffers, sampling_steps, step_coefficients  # noqa: E402

be = cabi.CudaBackend()
dev = "cuda"
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
B, C, S = 256, 3, 256                      # 16x the cfg2 batch: 201 MB per tensor
shape = (B, C, S, S)
bufs = bridge_buffers(1000, "linear", 1.0)
steps = sampling_steps(1000, True, "linear", 200)
coef = step_coefficients(bufs["m_t"], bufs["variance_t"], steps, 1.0)
x0, y, nz, eps = (torch.randn(shape, device=dev) for _ in range(4))
o1, o2 = torch.empty(shape, device=dev), torch.empty(shape, device=dev)
t = torch.randint(0, 1000, (B,), device=dev)
m_t, v_t = bufs["m_t"].to(dev), bufs["variance_t"].to(dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


n = x0.numel()
rows = []
ms = timeit(lambda: be.q_sample(x0, y, nz, t, m_t, v_t, "grad", o1, o2))
rows.append({"kernel": "q_sample_kernel<grad>", "bytes": 20 * n, "ms": ms})
ms = timeit(lambda: be.p_sample(x0, y, eps, nz, coef[10].tolist(), "grad", False, False, o1, o2))
rows.append({"kernel": "p_sample_kernel<grad,noclip,mid>", "bytes": 24 * n, "ms": ms})
ms = timeit(lambda: be.p_sample(x0, y, eps, None, coef[-1].tolist(), "grad", True, True, o1, None))
rows.append({"kernel": "p_sample_kernel<grad,clip,last>", "bytes": 16 * n, "ms": ms})
for r in rows:
    r["GBps"] = r["bytes"] / r["ms"] / 1e6
    r["frac_of_measured_hbm"] = r["GBps"] / peak
print(json.dumps({"tensor": list(shape), "hbm_peak_GBps": peak, "rows": rows}))
