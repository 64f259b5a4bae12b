"""Training path: the ResBlock convolutions (96 % of the training FLOPs) as an autograd Function
over the tcgen05 kernels -- forward (bbdm_conv_umma), data gradient (bbdm_conv_umma with the
flipped/transposed weights) and weight gradient (bbdm_conv_wgrad), all split-bf16 x3 with fp32
accumulation (fp32-class accuracy, like the reference's fp32 autograd).

Tensors cross the Function boundary as ordinary NCHW-shaped torch tensors in channels_last
memory format, i.e. physically the NHWC layout the kernels use: no layout copies when the
surrounding ops keep channels_last.  GroupNorm+SiLU+FiLM(+resampling) in front of a conv and the
attention core have their own Functions below (DESIGN.md "Training").

Replaces the autograd of nn.Conv2d inside ResBlock (reference openaimodel.py:207,233,244) for
``loss.backward()`` (runners/BaseRunner.py:412); gradients land in the same nn.Parameter.grad, so
DDP's bucketed NCCL allreduce works unchanged.
"""
from __future__ import annotations

import os

import torch

from . import cabi

_BACKEND = None

# Winograd F(4x4,3x3) for the forward and data-gradient 3x3 convolutions of the training graph (csrc/winograd.cu;
# same eligibility rule as the sampling engine).  The weight gradient stays the direct tcgen05 GEMM.
WINO_TRAIN = os.environ.get("BBDM_WINOGRAD_TRAIN", "1") != "0"
WINO_MIN_C = int(os.environ.get("BBDM_WINO_MIN_C", "256"))
WINO_MIN_TILES = int(os.environ.get("BBDM_WINO_MIN_TILES", "512"))


_WARNED = set()


def _library_path(what, x):
    """A CUDA training call whose shape the native kernels do not take runs on stock PyTorch (library) kernels: valid
    results, but not this library's path -- say so once per shape instead of falling back silently."""
    if x.is_cuda:
        key = (what, tuple(x.shape))
        if key not in _WARNED:
            _WARNED.add(key)
            import warnings
            warnings.warn(f"bbdm_b200.train: {what} with input {tuple(x.shape)} runs on stock PyTorch kernels "
                          "(shape outside the tensor-core kernels' envelope)", stacklevel=3)


def _wino_ok(be, B, H, W, Cin, Cout, k):
    if not WINO_TRAIN or k != 3 or min(Cin, Cout) < WINO_MIN_C or Cin % 64 or Cout % 64 or not hasattr(be, "wino_geometry"):
        return False
    th, tw, tiles, ok = be.wino_geometry(B, H, W)
    return bool(ok and (th * tw >= 128 or tiles >= WINO_MIN_TILES))


def _wino_conv(be, xn, weight, *, dgrad, gn=None, act_planes=None, bias=None, residual=None, out_channels):
    """3x3 conv of the NHWC tensor xn on the Winograd path: input transform (GroupNorm affine + FiLM + SiLU fused when
    gn = dict(mean, rstd, gamma, beta, film_scale, film_shift, film_stride, silu); identity for gn None), 36
    position GEMMs, output transform (+ bias + residual).  dgrad: use the flipped / channel-swapped kernel."""
    B, H, W, C = xn.shape
    dev = xn.device
    _, _, mt, _ = be.wino_geometry(B, H, W)
    v_hi = torch.empty((36, mt, C), dtype=torch.float16, device=dev)
    v_lo = torch.empty_like(v_hi)
    kw = dict(silu=False) if gn is None else gn
    akw = {} if act_planes is None else dict(act_hi=act_planes[0], act_lo=act_planes[1])
    be.wino_input(xn, None, v_hi=v_hi, v_lo=v_lo, **kw, **akw)
    u_hi = torch.empty((36, out_channels, C), dtype=torch.float16, device=dev)
    u_lo = torch.empty_like(u_hi)
    be.wino_pack_weight(weight.detach().contiguous(), u_hi, u_lo, dgrad=dgrad)
    m = torch.empty((36, mt, out_channels), dtype=torch.float32, device=dev)
    be.conv_umma(B=36, H=mt // 16, W=16, Cin=C, Cout=out_channels, taps=1, a_hi=v_hi, a_lo=v_lo, w_hi=u_hi, w_lo=u_lo,
                 out=m, passes=3, weights_per_image=True, operand_f16=True)
    out = torch.empty((B, H, W, out_channels), dtype=torch.float32, device=dev)
    be.wino_output(m, B=B, H=H, W=W, Cout=out_channels, bias=bias, residual=residual,
                   res_mode=cabi.RES_NONE if residual is None else cabi.RES_SAME, out=out)
    return out


def backend():
    global _BACKEND
    if _BACKEND is None:
        _BACKEND = cabi.CudaBackend()
    return _BACKEND


def set_backend(b):
    """tests/ inject their oracle-backed emulation here to check this module's host logic (gradient formulas,
    adjoints, tensor plumbing) on CPU; the product only ever uses cabi.CudaBackend."""
    global _BACKEND
    _BACKEND = b


def _on_device(x: torch.Tensor) -> bool:
    return x.is_cuda or (_BACKEND is not None and not getattr(_BACKEND, "requires_cuda", True))


def native_ok(conv: torch.nn.Conv2d, x: torch.Tensor) -> bool:
    """Shapes the tensor-core fwd/dgrad/wgrad kernels take."""
    if not _on_device(x) or x.dtype != torch.float32 or x.dim() != 4:
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


def _pack_weights(be, weight, need_dgrad):
    """(w_hi, w_lo, wd_hi, wd_lo): forward planes and, if the input needs a gradient, the data-gradient planes
    (flipped kernel, channels swapped) from the same single pass over the weight."""
    Cout, Cin, k, _ = weight.shape
    dev = weight.device
    w_hi = torch.empty((k * k, Cout, Cin), dtype=torch.bfloat16, device=dev)
    w_lo = torch.empty_like(w_hi)
    wd_hi = wd_lo = None
    if need_dgrad:
        wd_hi = torch.empty((k * k, Cin, Cout), dtype=torch.bfloat16, device=dev)
        wd_lo = torch.empty_like(wd_hi)
    be.pack_weight_split_both(weight.detach().contiguous(), w_hi, w_lo, wd_hi, wd_lo)
    return w_hi, w_lo, wd_hi, wd_lo


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
        w_hi, w_lo, wd_hi, wd_lo = _pack_weights(be, weight, ctx.needs_input_grad[0])
        out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=dev)
        be.conv_umma(B=B, H=H, W=W, Cin=Cin, Cout=Cout, taps=k * k, a_hi=a_hi, a_lo=a_lo, w_hi=w_hi, w_lo=w_lo,
                     bias=None if bias is None else bias.detach(), out=out, passes=3)
        ctx.save_for_backward(a_hi, a_lo, weight, wd_hi, wd_lo)
        ctx.has_bias = bias is not None
        ctx.shape = (B, H, W, Cin, Cout, k)
        return out.permute(0, 3, 1, 2)                                   # NCHW shape, channels_last strides

    @staticmethod
    def backward(ctx, dy):
        a_hi, a_lo, weight, wd_hi, wd_lo = ctx.saved_tensors
        dxn, dw, dbias = _conv_backward(backend(), ctx.shape, a_hi, a_lo, weight, dy, ctx.needs_input_grad[0],
                                        ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2],
                                        wd=(wd_hi, wd_lo))
        return (None if dxn is None else dxn.permute(0, 3, 1, 2)), dw, dbias


def _conv_backward(be, ctx_shape, a_hi, a_lo, weight, dy, need_dx, need_dw, need_db, wd=(None, None)):
    """Shared by both Functions: (dA or dX as NHWC fp32, dW, dbias) of the tensor-core conv."""
    B, H, W, Cin, Cout, k = ctx_shape
    dev = dy.device
    P = B * H * W
    dyn = _nhwc(dy)
    g_hi = g_lo = None
    wino_dx = need_dx and _wino_ok(be, B, H, W, Cout, Cin, k)
    if need_dx and not wino_dx:
        g_hi = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=dev)
        g_lo = torch.empty_like(g_hi)
    gt_hi = torch.empty((Cout, P), dtype=torch.bfloat16, device=dev)
    gt_lo = torch.empty_like(gt_hi)
    dbias = ws_b = None
    if need_db:
        dbias = torch.empty((Cout,), dtype=torch.float32, device=dev)
        ws_b = torch.empty(((P + 63) // 64) * Cout, dtype=torch.float32, device=dev)
    be.split_grad(dyn, g_hi, g_lo, gt_hi, gt_lo, dbias, ws_b)
    dxn = None
    if wino_dx:
        # data gradient = conv of dY with the flipped, channel-swapped kernel -- on the Winograd path.
        # Its operand planes are fp16 pairs (5 exponent bits): loss gradients (1e-4 ... 1e-8) would sit in fp16's
        # subnormal range and lose their mantissa (measured: 4e-3 per layer on the LBBDM-f4 UNet).  dY is therefore
        # normalised by a power of two that puts its largest element into [16, 32) -- the range the forward path's
        # activations live in -- and the result is scaled back; both scalings are exact, and the scale stays on the
        # device (no host synchronisation).
        amax = dyn.abs().amax().clamp_min(2.0 ** -100)
        scale = torch.exp2(4.0 - torch.floor(torch.log2(amax)))
        dxn = _wino_conv(be, (dyn * scale).contiguous(), weight, dgrad=True, out_channels=Cin)
        dxn.mul_(1.0 / scale)
    elif need_dx:
        # data gradient = the same conv with the kernel flipped and Cin/Cout swapped
        wd_hi, wd_lo = wd
        if wd_hi is None:
            wd_hi = torch.empty((k * k, Cin, Cout), dtype=torch.bfloat16, device=dev)
            wd_lo = torch.empty_like(wd_hi)
            be.pack_weight_split_dgrad(weight.detach().contiguous(), wd_hi, wd_lo)
        dxn = torch.empty((B, H, W, Cin), dtype=torch.float32, device=dev)
        be.conv_umma(B=B, H=H, W=W, Cin=Cout, Cout=Cin, taps=k * k, a_hi=g_hi, a_lo=g_lo, w_hi=wd_hi, w_lo=wd_lo,
                     out=dxn, passes=3)
    dw = None
    if need_dw:
        _, fl = be.wgrad_workspace(B, H, W, Cin, Cout, k * k)
        ws = torch.empty((fl,), dtype=torch.float32, device=dev)
        dw = torch.empty((Cout, Cin, k, k), dtype=torch.float32, device=dev)
        be.conv_wgrad(gt_hi, gt_lo, a_hi, a_lo, B, H, W, Cin, Cout, k * k, dw, ws)
    return dxn, dw, dbias


class GNActConv2dFn(torch.autograd.Function):
    """conv( silu( GroupNorm32(x) * (1 + scale) + shift ) ) fused: the forward is the sampling path's
    stats + prep + tcgen05 conv; the backward adds the two-pass GroupNorm/SiLU/FiLM gradient kernels."""

    @staticmethod
    def forward(ctx, x, gamma, beta, scale, shift, weight, bias, resample=0, residual=None, act=True):
        """resample: 0 none, 1 nearest-2x up, 2 2x2 average pool -- applied between SiLU and the conv
        (ResBlock(up/down), openaimodel.py:259-264)."""
        be = backend()
        B, Cin, Hs, Ws = x.shape
        H, W = (Hs * 2, Ws * 2) if resample == 1 else ((Hs // 2, Ws // 2) if resample == 2 else (Hs, Ws))
        Cout, _, k, _ = weight.shape
        dev = x.device
        xn = _nhwc(x.detach())
        mean = torch.empty((B, 32), dtype=torch.float32, device=dev)
        rstd = torch.empty_like(mean)
        ws = torch.empty((B * 32 * cabi.GN_MAX_SLICES * 2,), dtype=torch.float64, device=dev)
        be.gn_stats(xn, None, 32, 1e-5, mean, rstd, ws)
        fs = fh = None
        if scale is not None:
            fs, fh = scale.detach().contiguous().float(), shift.detach().contiguous().float()
        a_hi = torch.empty((B, H, W, Cin), dtype=torch.bfloat16, device=dev)
        a_lo = torch.empty_like(a_hi)
        rn = None if residual is None else _nhwc(residual.detach())       # + skip(x), fused in the epilogue
        if resample == 0 and _wino_ok(be, B, H, W, Cin, Cout, k):
            # Winograd forward; the input transform also writes the activated split planes the weight gradient needs
            gn = dict(groups=32, mean=mean, rstd=rstd, gamma=gamma.detach(), beta=beta.detach(), film_scale=fs,
                      film_shift=fh, film_stride=0 if fs is None else fs.shape[1], silu=act)
            out = _wino_conv(be, xn.contiguous(), weight, dgrad=False, gn=gn, act_planes=(a_hi, a_lo),
                             bias=None if bias is None else bias.detach(), residual=None if rn is None else rn.contiguous(),
                             out_channels=Cout)
            wd_hi = wd_lo = None              # the backward re-derives what it needs (Winograd dgrad planes)
        else:
            be.prep(xn, None, groups=32, mean=mean, rstd=rstd, gamma=gamma.detach(), beta=beta.detach(), film_scale=fs,
                    film_shift=fh, film_stride=0 if fs is None else fs.shape[1], silu=act, resample=resample,
                    act_hi=a_hi, act_lo=a_lo)
            w_hi, w_lo, wd_hi, wd_lo = _pack_weights(be, weight, True)
            out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=dev)
            be.conv_umma(B=B, H=H, W=W, Cin=Cin, Cout=Cout, taps=k * k, a_hi=a_hi, a_lo=a_lo, w_hi=w_hi, w_lo=w_lo,
                         bias=None if bias is None else bias.detach(), residual=rn,
                         res_mode=cabi.RES_NONE if rn is None else cabi.RES_SAME, out=out, passes=3)
        ctx.save_for_backward(xn, mean, rstd, gamma, beta, fs, fh, a_hi, a_lo, weight, wd_hi, wd_lo)
        ctx.has_bias = bias is not None
        ctx.shape = (B, H, W, Cin, Cout, k)
        ctx.resample = resample
        ctx.has_res = residual is not None
        ctx.act = act
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        be = backend()
        xn, mean, rstd, gamma, beta, fs, fh, a_hi, a_lo, weight, wd_hi, wd_lo = ctx.saved_tensors
        B, H, W, Cin, Cout, k = ctx.shape
        dev = dy.device
        da, dw, dbias = _conv_backward(be, ctx.shape, a_hi, a_lo, weight, dy, True, ctx.needs_input_grad[5],
                                       ctx.has_bias and ctx.needs_input_grad[6], wd=(wd_hi, wd_lo))
        if ctx.resample == 1:        # adjoint of nearest-2x: sum the four children
            da = da.view(B, H // 2, 2, W // 2, 2, Cin).sum(dim=(2, 4)).contiguous()
        elif ctx.resample == 2:      # adjoint of the 2x2 mean: a quarter to each of the four parents
            da = (0.25 * da).repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).contiguous()
        H, W = xn.shape[1], xn.shape[2]
        g, b_ = gamma.detach(), beta.detach()
        fstride = 0 if fs is None else fs.shape[1]
        a12 = torch.empty((B, Cin, 2), dtype=torch.float32, device=dev)
        ws = torch.empty((B * 64 * Cin * 2,), dtype=torch.float32, device=dev)
        be.gn_bwd_reduce(xn, da, 32, mean, rstd, g, b_, fs, fh, fstride, ctx.act, a12, ws)
        a1, a2 = a12[..., 0], a12[..., 1]                                  # [B, C]
        f1 = (1.0 + fs) if fs is not None else torch.ones_like(a1)
        dgamma = (f1 * a2).sum(0)
        dbeta = (f1 * a1).sum(0)
        dscale = dshift = None
        if fs is not None:
            dshift = a1
            dscale = g * a2 + b_ * a1
        gf = g * f1                                                       # [B, C]
        s1 = (gf * a1).view(B, 32, Cin // 32).sum(2).contiguous()
        s2 = (gf * a2).view(B, 32, Cin // 32).sum(2).contiguous()
        dxn = torch.empty((B, H, W, Cin), dtype=torch.float32, device=dev)
        be.gn_bwd_apply(xn, da, 32, mean, rstd, g, b_, fs, fh, fstride, ctx.act, s1, s2, dxn)
        return dxn.permute(0, 3, 1, 2), dgamma, dbeta, dscale, dshift, dw, dbias, None, (dy if ctx.has_res else None), None


def gn_act_conv2d(norm, conv, x, scale=None, shift=None, enabled=True, resample=0, residual=None, act=True):
    """conv(resample(silu(norm(x) * (1 + scale) + shift))) [+ residual] -- fused tensor-core path when the
    shape qualifies (the residual add then happens in the conv epilogue)."""
    B, _, Hs, Ws = x.shape
    H, W = (Hs * 2, Ws * 2) if resample == 1 else ((Hs // 2, Ws // 2) if resample == 2 else (Hs, Ws))
    probe = x if resample == 0 else x.new_empty((B, x.shape[1], H, W))       # shape check at the conv's resolution
    if enabled and native_ok(conv, probe) and x.shape[1] % 32 == 0 and x.shape[1] <= 4096 and \
            (resample != 2 or (Hs % 2 == 0 and Ws % 2 == 0)):
        sc = None if scale is None else scale.reshape(scale.shape[0], -1)
        sh = None if shift is None else shift.reshape(shift.shape[0], -1)
        return GNActConv2dFn.apply(x, norm.weight, norm.bias, sc, sh, conv.weight, conv.bias, resample, residual, act)
    if enabled:
        _library_path("GroupNorm+activation+conv", x)
    h = norm(x)
    if scale is not None:
        h = h * (1 + scale) + shift
    if act:
        h = torch.nn.functional.silu(h)
    if resample == 1:
        h = torch.nn.functional.interpolate(h, scale_factor=2, mode="nearest")
    elif resample == 2:
        h = torch.nn.functional.avg_pool2d(h, 2)
    h = conv2d(conv, h, enabled)
    return h if residual is None else residual + h


class SmallConv2dFn(torch.autograd.Function):
    """The UNet stem / head (3..32 channels on one side): exact fp32 on CUDA cores -- forward and data
    gradient through bbdm_conv_direct, weight gradient through bbdm_conv_wgrad_direct."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        be = backend()
        B, Cin, H, W = x.shape
        Cout, _, k, _ = weight.shape
        dev = x.device
        xn = _nhwc(x.detach()).contiguous()
        wp = torch.empty((k * k, Cin, Cout), dtype=torch.float32, device=dev)
        be.pack_weight_f32(weight.detach().contiguous(), wp)
        out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=dev)
        be.conv_direct(xn, wp, None if bias is None else bias.detach(), None, out, Cout, k, 1)
        ctx.save_for_backward(xn, weight)
        ctx.has_bias = bias is not None
        ctx.k = k
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        be = backend()
        xn, weight = ctx.saved_tensors
        k = ctx.k
        B, H, W, Cin = xn.shape
        Cout = weight.shape[0]
        dev = dy.device
        dyn = _nhwc(dy).contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wd = weight.detach().flip(2, 3).transpose(0, 1).contiguous()          # [Cin, Cout, k, k]
            wdp = torch.empty((k * k, Cout, Cin), dtype=torch.float32, device=dev)
            be.pack_weight_f32(wd, wdp)
            dxn = torch.empty((B, H, W, Cin), dtype=torch.float32, device=dev)
            be.conv_direct(dyn, wdp, None, None, dxn, Cin, k, 1)
            dx = dxn.permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            n = k * k * Cin * Cout
            ws = torch.empty((2048 * n,), dtype=torch.float32, device=dev)
            dw = torch.empty((Cout, Cin, k, k), dtype=torch.float32, device=dev)
            be.conv_wgrad_direct(dyn, xn, k, dw, ws)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dyn.sum(dim=(0, 1, 2))
        return dx, dw, db


def small_ok(conv: torch.nn.Conv2d, x: torch.Tensor) -> bool:
    k = conv.kernel_size
    return (_on_device(x) and x.dtype == torch.float32 and x.dim() == 4 and k in ((1, 1), (3, 3))
            and conv.stride == (1, 1) and conv.padding == (k[0] // 2, k[0] // 2) and conv.groups == 1
            and conv.dilation == (1, 1) and conv.in_channels * conv.out_channels <= 1024
            and min(conv.in_channels, conv.out_channels) <= 32)


def conv1x1(conv1d: torch.nn.Conv1d, x4: torch.Tensor, enabled: bool = True):
    """nn.Conv1d(k=1) of AttentionBlock (qkv / proj_out, openaimodel.py:307,315) applied to a [B,C,H,W]
    tensor as a 1x1 convolution on the tensor-core autograd path; returns [B,Cout,H,W] or None if the
    shape does not qualify (caller falls back to the module)."""
    B, C, H, W = x4.shape
    Cout = conv1d.out_channels
    ok = (enabled and _on_device(x4) and x4.dtype == torch.float32 and conv1d.kernel_size == (1,) and C % 64 == 0
          and Cout % 64 == 0 and W >= 4 and (B * H * W) % 64 == 0 and _box64_ok(B, H, W))
    if not ok:
        return None
    return Conv2dFn.apply(x4, conv1d.weight.unsqueeze(-1), conv1d.bias)


class AttentionCoreFn(torch.autograd.Function):
    """softmax((q s)(k s)^T) v per head (QKVAttentionLegacy / QKVAttention, openaimodel.py:350-413) on a
    [B,3C,H,W] qkv tensor -> [B,C,H,W].  Forward: the sampling path's attention kernels (tcgen05 for
    head_dim 64); backward: bbdm_attention_bwd (flash-style recompute, exact fp32) -- the T x T matrix is
    never stored, which also replaces the reference's checkpoint() around the block (openaimodel.py:318)."""

    @staticmethod
    def forward(ctx, qkv, heads, order):
        be = backend()
        B, C3, H, W = qkv.shape
        Cc, T = C3 // 3, H * W
        dev = qkv.device
        qn = _nhwc(qkv.detach()).contiguous()
        out = torch.empty((B, T, Cc), dtype=torch.float32, device=dev)
        if Cc // heads == 64:
            q_hi = torch.empty((B, H, W, C3), dtype=torch.bfloat16, device=dev)
            q_lo = torch.empty_like(q_hi)
            be.prep(qn, None, raw_hi=q_hi, raw_lo=q_lo)
            be.attention_tc(q_hi.view(B, T, C3), q_lo.view(B, T, C3), heads, order, out, None, None)
        else:
            be.attention(qn.view(B, T, C3), heads, order, out, None, None)
        ctx.save_for_backward(qn, out)
        ctx.heads, ctx.order = heads, order
        return out.view(B, H, W, Cc).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dout):
        be = backend()
        qn, out = ctx.saved_tensors
        B, H, W, C3 = qn.shape
        T = H * W
        dev = dout.device
        don = _nhwc(dout).contiguous()
        dqkv = torch.empty_like(qn)
        lse = torch.empty((B * ctx.heads * T,), dtype=torch.float32, device=dev)
        delta = torch.empty_like(lse)
        be.attention_bwd(qn.view(B, T, C3), out, don.view(B, T, C3 // 3), ctx.heads, ctx.order,
                         dqkv.view(B, T, C3), lse, delta)
        return dqkv.permute(0, 3, 1, 2), None, None


def attention_core(qkv4: torch.Tensor, heads: int, new_order: bool, enabled: bool = True):
    """[B,3C,H,W] -> [B,C,H,W] or None when the native kernels do not take the shape."""
    B, C3, H, W = qkv4.shape
    hd = C3 // 3 // heads
    if not (enabled and _on_device(qkv4) and qkv4.dtype == torch.float32 and hd in (16, 32, 64) and (C3 // 3) % 4 == 0
            and B * heads <= 65535):
        return None
    return AttentionCoreFn.apply(qkv4, heads, 1 if new_order else 0)


def gn_conv1x1(norm, conv1d: torch.nn.Conv1d, x4: torch.Tensor, enabled: bool = True):
    """conv1d_k1(GroupNorm32(x)) of AttentionBlock (openaimodel.py:307,321) fused like gn_act_conv2d, without
    the activation; None if the shape does not qualify."""
    B, C, H, W = x4.shape
    Cout = conv1d.out_channels
    ok = (enabled and _on_device(x4) and x4.dtype == torch.float32 and conv1d.kernel_size == (1,) and C % 64 == 0
          and Cout % 64 == 0 and W >= 4 and (B * H * W) % 64 == 0 and _box64_ok(B, H, W) and C <= 4096)
    if not ok:
        return None
    return GNActConv2dFn.apply(x4, norm.weight, norm.bias, None, None, conv1d.weight.unsqueeze(-1), conv1d.bias,
                               0, None, False)


def conv2d(conv: torch.nn.Conv2d, x: torch.Tensor, enabled: bool = True) -> torch.Tensor:
    """nn.Conv2d call with the tensor-core autograd path when the shape qualifies."""
    if enabled and native_ok(conv, x):
        return Conv2dFn.apply(x, conv.weight, conv.bias)
    if enabled and small_ok(conv, x):
        return SmallConv2dFn.apply(x, conv.weight, conv.bias)
    if enabled:
        _library_path(f"Conv2d {conv.in_channels}->{conv.out_channels} k{conv.kernel_size[0]}", x)
    return conv(x)
