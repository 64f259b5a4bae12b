"""Training path: the ResBlock convolutions (96 % of the training FLOPs) as an autograd Function
over the tcgen05 kernels -- forward (bbdm_conv_umma), data gradient (bbdm_conv_umma with the
flipped/transposed weights) and weight gradient (bbdm_conv_wgrad), all split-bf16 x3 with fp32
accumulation (fp32-class accuracy, like the reference's fp32 autograd).

Tensors cross the Function boundary as ordinary NCHW-shaped torch tensors in channels_last
memory format, i.e. physically the NHWC layout the kernels use: no layout copies when the
surrounding ops keep channels_last.  GroupNorm / SiLU / attention / resampling stay on PyTorch
autograd in this round (DESIGN.md "Training").

Replaces the autograd of nn.Conv2d inside ResBlock (reference openaimodel.py:207,233,244) for
``loss.backward()`` (runners/BaseRunner.py:412); gradients land in the same nn.Parameter.grad, so
DDP's bucketed NCCL allreduce works unchanged.
"""
from __future__ import annotations

import torch

from . import cabi

_BACKEND = None


def backend():
    global _BACKEND
    if _BACKEND is None:
        _BACKEND = cabi.CudaBackend()
    return _BACKEND


def native_ok(conv: torch.nn.Conv2d, x: torch.Tensor) -> bool:
    """Shapes the tensor-core fwd/dgrad/wgrad kernels take."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4:
        return False
    k = conv.kernel_size
    B, _, H, W = x.shape
    return (k in ((1, 1), (3, 3)) and conv.stride == (1, 1) and conv.padding == (k[0] // 2, k[0] // 2)
            and conv.groups == 1 and conv.dilation == (1, 1)
            and conv.in_channels % 64 == 0 and conv.out_channels % 64 == 0 and W >= 4
            and (B * H * W) % 64 == 0 and _box64_ok(B, H, W))


def _box64_ok(B, H, W):
    tw = 1
    while tw * 2 <= W and tw * 2 <= 64 and W % (tw * 2) == 0:
        tw *= 2
    th = 1
    while tw * th * 2 <= 64 and th * 2 <= H and H % (th * 2) == 0:
        th *= 2
    tb = 64 // (tw * th)
    return W % tw == 0 and H % th == 0 and (tw == W or th == 1) and (th == H or tb == 1) and B % tb == 0


def _nhwc(x):
    """NCHW-shaped tensor -> contiguous NHWC view (copy only if x is not channels_last already)."""
    return x.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)


class Conv2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        be = backend()
        B, Cin, H, W = x.shape
        Cout, _, k, _ = weight.shape
        dev = x.device
        xn = _nhwc(x.detach())
        a_hi = torch.empty((B, H, W, Cin), dtype=torch.bfloat16, device=dev)
        a_lo = torch.empty_like(a_hi)
        be.prep(xn, None, raw_hi=a_hi, raw_lo=a_lo)                      # operand split (one HBM pass)
        w_hi = torch.empty((k * k, Cout, Cin), dtype=torch.bfloat16, device=dev)
        w_lo = torch.empty_like(w_hi)
        be.pack_weight_split(weight.detach().contiguous(), w_hi, w_lo)
        out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=dev)
        be.conv_umma(B=B, H=H, W=W, Cin=Cin, Cout=Cout, taps=k * k, a_hi=a_hi, a_lo=a_lo, w_hi=w_hi, w_lo=w_lo,
                     bias=None if bias is None else bias.detach(), out=out, passes=3)
        ctx.save_for_backward(a_hi, a_lo, weight)
        ctx.has_bias = bias is not None
        ctx.shape = (B, H, W, Cin, Cout, k)
        return out.permute(0, 3, 1, 2)                                   # NCHW shape, channels_last strides

    @staticmethod
    def backward(ctx, dy):
        be = backend()
        a_hi, a_lo, weight = ctx.saved_tensors
        B, H, W, Cin, Cout, k = ctx.shape
        dev = dy.device
        P = B * H * W
        dyn = _nhwc(dy)
        need_dx = ctx.needs_input_grad[0]
        g_hi = g_lo = None
        if need_dx:
            g_hi = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=dev)
            g_lo = torch.empty_like(g_hi)
        gt_hi = torch.empty((Cout, P), dtype=torch.bfloat16, device=dev)
        gt_lo = torch.empty_like(gt_hi)
        dbias = ws_b = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            dbias = torch.empty((Cout,), dtype=torch.float32, device=dev)
            ws_b = torch.empty(((P + 63) // 64) * Cout, dtype=torch.float32, device=dev)
        be.split_grad(dyn, g_hi, g_lo, gt_hi, gt_lo, dbias, ws_b)

        dx = None
        if need_dx:
            # data gradient = the same conv with the kernel flipped and Cin/Cout swapped
            wd = weight.detach().flip(2, 3).transpose(0, 1).contiguous()          # [Cin, Cout, k, k]
            wd_hi = torch.empty((k * k, Cin, Cout), dtype=torch.bfloat16, device=dev)
            wd_lo = torch.empty_like(wd_hi)
            be.pack_weight_split(wd, wd_hi, wd_lo)
            dxn = torch.empty((B, H, W, Cin), dtype=torch.float32, device=dev)
            be.conv_umma(B=B, H=H, W=W, Cin=Cout, Cout=Cin, taps=k * k, a_hi=g_hi, a_lo=g_lo, w_hi=wd_hi, w_lo=wd_lo,
                         out=dxn, passes=3)
            dx = dxn.permute(0, 3, 1, 2)
        dw = None
        if ctx.needs_input_grad[1]:
            _, fl = be.wgrad_workspace(B, H, W, Cin, Cout, k * k)
            ws = torch.empty((fl,), dtype=torch.float32, device=dev)
            dw = torch.empty((Cout, Cin, k, k), dtype=torch.float32, device=dev)
            be.conv_wgrad(gt_hi, gt_lo, a_hi, a_lo, B, H, W, Cin, Cout, k * k, dw, ws)
        return dx, dw, dbias


def conv2d(conv: torch.nn.Conv2d, x: torch.Tensor, enabled: bool = True) -> torch.Tensor:
    """nn.Conv2d call with the tensor-core autograd path when the shape qualifies."""
    if enabled and native_ok(conv, x):
        return Conv2dFn.apply(x, conv.weight, conv.bias)
    return conv(x)
