"""Brownian-bridge schedule: the six fp32 [T] buffers, the sampling step list and the per-step
scalar coefficients of the reverse update.

Restates BrownianBridgeModel.register_schedule (reference
model/BrownianBridge/BrownianBridgeModel.py:42-79).  The buffers are computed in float64 numpy
and cast to fp32 once, as in the reference, so they match it bit for bit
(tests/golden/schedule_kats.json holds reference-generated hashes).
"""
from __future__ import annotations

import numpy as np
import torch

BUFFER_NAMES = ("m_t", "m_tminus", "variance_t", "variance_tminus", "variance_t_tminus",
                "posterior_variance_t")


def bridge_buffers(num_timesteps: int, mt_type: str, max_var: float) -> dict:
    T = num_timesteps
    if mt_type == "linear":
        m = np.linspace(0.001, 0.999, T)
    elif mt_type == "sin":
        m = 1.0075 ** np.linspace(0, T, T)
        m = m / m[-1]
        m[-1] = 0.999
    else:
        raise NotImplementedError
    var = 2.0 * (m - m ** 2) * max_var
    m_prev = np.append(0, m[:-1])
    var_prev = np.append(0.0, var[:-1])
    var_step = var - var_prev * ((1.0 - m) / (1.0 - m_prev)) ** 2
    post = var_step * var_prev / var
    vals = (m, m_prev, var, var_prev, var_step, post)
    return {k: torch.tensor(v, dtype=torch.float32) for k, v in zip(BUFFER_NAMES, vals)}


def sampling_steps(num_timesteps: int, skip_sample: bool, sample_type: str, sample_step: int):
    """CPU int64 tensor of timesteps visited by p_sample_loop (reference :68-79)."""
    T = num_timesteps
    if not skip_sample:
        return torch.arange(T - 1, -1, -1)
    if sample_type == "linear":
        mid = torch.arange(T - 1, 1, step=-((T - 1) / (sample_step - 2))).long()
        return torch.cat((mid, torch.Tensor([1, 0]).long()), dim=0)
    if sample_type == "cosine":
        # Faithful to the reference, including its defect: a float64 list starting at
        # T, which makes the first gather index out of range there too.
        s = np.linspace(start=0, stop=T, num=sample_step + 1)
        return torch.from_numpy((np.cos(s / T * np.pi) + 1.0) / 2.0 * T)
    return None     # the reference leaves self.steps = None for unknown sample_type


def step_coefficients(m_t: torch.Tensor, variance_t: torch.Tensor, steps: torch.Tensor, eta):
    """[n_steps, 7] fp32 table (m_t, 1-m_t, sqrt(var_t), m_nt, 1-m_nt, c_xt, sigma_t) per step.

    Evaluated with fp32 torch ops in the reference's exact expression order
    (BrownianBridgeModel.py:190-199) so the fused kernel reproduces its scalars bit for bit.
    The last row (t == 0) only uses the first three entries.
    """
    m_t = m_t.detach().float().cpu()
    variance_t = variance_t.detach().float().cpu()
    st = steps.long()
    n = len(st)
    t = st
    nt = torch.cat([st[1:], st[-1:]])            # next step (dummy for the final row)
    m, mn = m_t[t], m_t[nt]
    v, vn = variance_t[t], variance_t[nt]
    sigma2 = (v - vn * (1. - m) ** 2 / (1. - mn) ** 2) * vn / v
    sigma = torch.sqrt(sigma2) * eta
    c_xt = torch.sqrt((vn - sigma2) / v)
    tab = torch.stack([m, 1. - m, torch.sqrt(v), mn, 1. - mn, c_xt, sigma], dim=1).float()
    last = (st == 0)
    tab[last, 3:] = 0.0
    assert tab.shape == (n, 7)
    return tab.contiguous()
