#!/usr/bin/env python
"""CPU study (no GPU): would Winograd F(2x2,3x3) keep the <=1e-4 parity budget with split-bf16 x3 operands?

Emulates, in fp64 arithmetic on exactly-representable bf16 planes, what a tcgen05 implementation would compute:
  direct  : A_hi*W_hi + A_lo*W_hi + A_hi*W_lo                                   (today's kernel)
  winograd: V = B^T d B (fp32) -> split; U = G g G^T (fp32, from fp64) -> split; M = V_hi*U_hi + V_lo*U_hi + V_hi*U_lo;
            Y = A^T M A in fp32.
and compares both with the fp64 direct convolution, for activation / weight statistics of the UNet
(post-GroupNorm+SiLU activations, N(0, 0.02)-class weights) at several channel counts.
"""
import json
import sys

import torch
import torch.nn.functional as F

torch.manual_seed(0)
BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float64)
G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float64)
AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float64)


def split(x):
    x = x.float()
    hi = x.to(torch.bfloat16).float()
    lo = (x - hi).to(torch.bfloat16).float()
    return hi.double(), lo.double()


def direct_split3(x, w):
    xh, xl = split(x)
    wh, wl = split(w)
    return F.conv2d(xh, wh, padding=1) + F.conv2d(xl, wh, padding=1) + F.conv2d(xh, wl, padding=1)


def winograd_split3(x, w):
    B, C, H, W = x.shape
    K = w.shape[0]
    xp = F.pad(x.double(), (1, 1, 1, 1))
    # 4x4 tiles, stride 2: [B, C, th, tw, 4, 4]
    t = xp.unfold(2, 4, 2).unfold(3, 4, 2)
    V = torch.einsum("ij,bcxyjk,lk->bcxyil", BT, t, BT).float()          # input transform, rounded to fp32
    U = torch.einsum("ij,kcjl,ml->kcim", G, w.double(), G).float()        # weight transform, rounded to fp32
    Vh, Vl = split(V)
    Uh, Ul = split(U)
    M = sum(torch.einsum("bcxyil,kcil->bkxyil", a, b) for a, b in ((Vh, Uh), (Vl, Uh), (Vh, Ul))).float()
    Y = torch.einsum("ij,bkxyjl,ml->bkxyim", AT, M.double(), AT).float()  # output transform in fp32
    th, tw = Y.shape[2], Y.shape[3]
    return Y.permute(0, 1, 2, 4, 3, 5).reshape(B, K, th * 2, tw * 2).double()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def main():
    rows = []
    for C, K, HW in ((64, 64, 16), (256, 256, 16), (512, 512, 8), (1024, 512, 8)):
        x = F.silu(torch.randn(1, C, HW, HW) * 1.0 + 0.1)                 # post GN+SiLU-like
        w = torch.randn(K, C, 3, 3) * 0.02
        ref = F.conv2d(x.double(), w.double(), padding=1)
        d, wg = direct_split3(x, w), winograd_split3(x, w)
        # sanity of the transform itself in fp64
        t = F.pad(x.double(), (1, 1, 1, 1)).unfold(2, 4, 2).unfold(3, 4, 2)
        V = torch.einsum("ij,bcxyjk,lk->bcxyil", BT, t, BT)
        U = torch.einsum("ij,kcjl,ml->kcim", G, w.double(), G)
        Y = torch.einsum("ij,bkxyjl,ml->bkxyim", AT, torch.einsum("bcxyil,kcil->bkxyil", V, U), AT)
        exact = Y.permute(0, 1, 2, 4, 3, 5).reshape(ref.shape)
        rows.append({"Cin": C, "Cout": K, "HW": HW, "direct_split3_rel_dev": rel(d, ref), "winograd_split3_rel_dev": rel(wg, ref),
                     "ratio": rel(wg, ref) / rel(d, ref), "winograd_fp64_rel_dev": rel(exact, ref)})
        print(json.dumps(rows[-1]))
    return rows


if __name__ == "__main__":
    main()
