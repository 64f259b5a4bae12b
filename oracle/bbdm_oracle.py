"""CPU oracle for the BBDM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``bbdm_b200``) never imports it and has no CPU fallback.

This is an independent *functional* restatement (plain torch CPU ops on a
``state_dict``; no nn.Module tree) of the reference algorithm for the path
BASELINE.json's north_star names:

* Brownian-bridge schedule, q_sample, p_sample, p_sample_loop
  -> /root/reference/model/BrownianBridge/BrownianBridgeModel.py:42-225
* the denoising UNet forward
  -> /root/reference/model/BrownianBridge/base/modules/diffusionmodules/openaimodel.py:166-327,416-759
  -> /root/reference/model/BrownianBridge/base/modules/diffusionmodules/util.py:151-216

Parity pinning: ``tests/golden/make_golden.py`` imports the *unmodified reference*
from /root/reference in the build container, runs it on seeded inputs and commits
the outputs under ``tests/golden/``; ``tests/test_oracle_pinned.py`` checks this
restatement against those fixtures (and, when /root/reference is present, against
the live reference).  So parity is PINNED (reference-generated fixtures).

Everything is computed in the dtype of the inputs (fp32 like the reference, or
fp64 when the caller up-casts the state dict to get a "true value" to measure
both implementations against).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Schedule  (BrownianBridgeModel.py:42-79)
# --------------------------------------------------------------------------------------


def make_schedule(num_timesteps=1000, mt_type="linear", max_var=1.0, skip_sample=True,
                  sample_type="linear", sample_step=200):
    """Returns (dict of six fp32 [T] tensors, steps int64 tensor).

    float64 numpy math then a cast to fp32, exactly like register_schedule
    (BrownianBridgeModel.py:42-66); ``steps`` as :68-79.
    """
    T = num_timesteps
    if mt_type == "linear":
        m_t = np.linspace(0.001, 0.999, T)
    elif mt_type == "sin":
        m_t = 1.0075 ** np.linspace(0, T, T)
        m_t = m_t / m_t[-1]
        m_t[-1] = 0.999
    else:
        raise NotImplementedError
    m_tminus = np.append(0, m_t[:-1])
    variance_t = 2.0 * (m_t - m_t ** 2) * max_var
    variance_tminus = np.append(0.0, variance_t[:-1])
    variance_t_tminus = variance_t - variance_tminus * ((1.0 - m_t) / (1.0 - m_tminus)) ** 2
    posterior_variance_t = variance_t_tminus * variance_tminus / variance_t
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    bufs = dict(m_t=f32(m_t), m_tminus=f32(m_tminus), variance_t=f32(variance_t),
                variance_tminus=f32(variance_tminus), variance_t_tminus=f32(variance_t_tminus),
                posterior_variance_t=f32(posterior_variance_t))
    if skip_sample:
        if sample_type == "linear":
            mid = torch.arange(T - 1, 1, step=-((T - 1) / (sample_step - 2))).long()
            steps = torch.cat((mid, torch.tensor([1, 0], dtype=torch.long)), dim=0)
        else:
            # 'cosine' is broken in the reference; not restated.
            raise NotImplementedError(sample_type)
    else:
        steps = torch.arange(T - 1, -1, -1)
    return bufs, steps


def _extract(a, t, ndim):
    # model/utils.py:4-7
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (ndim - 1)))


# --------------------------------------------------------------------------------------
# Bridge elementwise  (BrownianBridgeModel.py:128-160, 171-201)
# --------------------------------------------------------------------------------------


def q_sample(bufs, x0, y, t, noise, objective="grad"):
    m_t = _extract(bufs["m_t"].to(x0.dtype), t, x0.dim())
    var_t = _extract(bufs["variance_t"].to(x0.dtype), t, x0.dim())
    sigma_t = torch.sqrt(var_t)
    if objective == "grad":
        obj = m_t * (y - x0) + sigma_t * noise
    elif objective == "noise":
        obj = noise
    elif objective == "ysubx":
        obj = y - x0
    else:
        raise NotImplementedError
    return (1.0 - m_t) * x0 + m_t * y + sigma_t * noise, obj


def predict_x0(bufs, x_t, y, t, eps, objective="grad"):
    if objective == "grad":
        return x_t - eps
    if objective == "noise":
        m_t = _extract(bufs["m_t"].to(x_t.dtype), t, x_t.dim())
        var_t = _extract(bufs["variance_t"].to(x_t.dtype), t, x_t.dim())
        return (x_t - m_t * y - torch.sqrt(var_t) * eps) / (1.0 - m_t)
    if objective == "ysubx":
        return y - eps
    raise NotImplementedError


def p_sample_update(bufs, steps, i, x_t, y, eps, noise, objective="grad", eta=1.0,
                    clip_denoised=False):
    """The elementwise part of p_sample (everything after the UNet call).

    Returns (x_{t-1}, x0_recon).  BrownianBridgeModel.py:174-201.
    """
    B = x_t.shape[0]
    t = torch.full((B,), int(steps[i]), dtype=torch.long)
    x0 = predict_x0(bufs, x_t, y, t, eps, objective)
    if clip_denoised:
        x0 = x0.clamp(-1.0, 1.0)
    if int(steps[i]) == 0:
        return x0, x0
    n_t = torch.full((B,), int(steps[i + 1]), dtype=torch.long)
    dt = x_t.dtype
    m_t = _extract(bufs["m_t"].to(dt), t, x_t.dim())
    m_nt = _extract(bufs["m_t"].to(dt), n_t, x_t.dim())
    var_t = _extract(bufs["variance_t"].to(dt), t, x_t.dim())
    var_nt = _extract(bufs["variance_t"].to(dt), n_t, x_t.dim())
    sigma2_t = (var_t - var_nt * (1.0 - m_t) ** 2 / (1.0 - m_nt) ** 2) * var_nt / var_t
    sigma_t = torch.sqrt(sigma2_t) * eta
    mean = (1.0 - m_nt) * x0 + m_nt * y + torch.sqrt((var_nt - sigma2_t) / var_t) * \
        (x_t - (1.0 - m_t) * x0 - m_t * y)
    return mean + sigma_t * noise, x0


# --------------------------------------------------------------------------------------
# UNet structure  (openaimodel.py:446-703) -- re-derived here, independently of the product
# --------------------------------------------------------------------------------------

DEFAULT_UNET = dict(dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                    num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=-1,
                    num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                    resblock_updown=False, use_new_attention_order=False,
                    use_spatial_transformer=False, transformer_depth=1, context_dim=None,
                    n_embed=None, legacy=True, condition_key="concat")


def unet_cfg(**kw):
    d = dict(DEFAULT_UNET)
    d.update(kw)
    return SimpleNamespace(**d)


def _block_plan(cfg):
    """List of blocks as (prefix, [layer descriptors]) following UNetModel.__init__."""
    mc = cfg.model_channels
    plan = {"input": [], "middle": [], "output": []}
    plan["input"].append([("conv", None)])
    chans = [mc]
    ch, ds = mc, 1
    nh = cfg.num_heads

    def heads_for(c, nh_in):
        if cfg.num_head_channels == -1:
            return nh_in
        return c // cfg.num_head_channels

    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", dict(cin=ch, cout=mult * mc, up=False, down=False))]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers.append(("attn", dict(c=ch, heads=heads_for(ch, nh))))
            plan["input"].append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            if cfg.resblock_updown:
                plan["input"].append([("res", dict(cin=ch, cout=ch, up=False, down=True))])
            else:
                plan["input"].append([("downsample", dict(c=ch, use_conv=cfg.conv_resample))])
            chans.append(ch)
            ds *= 2
    plan["middle"] = [("res", dict(cin=ch, cout=ch, up=False, down=False)),
                      ("attn", dict(c=ch, heads=heads_for(ch, nh))),
                      ("res", dict(cin=ch, cout=ch, up=False, down=False))]
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            layers = [("res", dict(cin=ch + ich, cout=mc * mult, up=False, down=False))]
            ch = mc * mult
            if ds in cfg.attention_resolutions:
                layers.append(("attn", dict(c=ch, heads=heads_for(ch, cfg.num_heads_upsample
                                                               if cfg.num_heads_upsample != -1 else nh))))
            if level and i == cfg.num_res_blocks:
                if cfg.resblock_updown:
                    layers.append(("res", dict(cin=ch, cout=ch, up=True, down=False)))
                else:
                    layers.append(("upsample", dict(c=ch, use_conv=cfg.conv_resample)))
                ds //= 2
            plan["output"].append(layers)
    return plan


def timestep_embedding(t, dim, max_period=10000):
    # util.py:151-171 (cos first, then sin)
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(x, sd, p, eps=1e-5):
    # util.py:199-216: GroupNorm(32, C), eps 1e-5
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _resblock(sd, p, x, emb, d, use_scale_shift_norm):
    # openaimodel.py:258-278
    h = F.silu(_gn(x, sd, p + ".in_layers.0"))
    if d["up"]:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif d["down"]:
        h = F.avg_pool2d(h, 2)
        x = F.avg_pool2d(x, 2)
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    e = e[:, :, None, None]
    if use_scale_shift_norm:
        scale, shift = torch.chunk(e, 2, dim=1)
        h = _gn(h, sd, p + ".out_layers.0") * (1 + scale) + shift
        h = F.silu(h)
    else:
        h = h + e
        h = F.silu(_gn(h, sd, p + ".out_layers.0"))
    h = F.conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if p + ".skip_connection.weight" in sd:
        w = sd[p + ".skip_connection.weight"]
        x = F.conv2d(x, w, sd[p + ".skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


def qkv_attention(qkv, heads, new_order=False):
    """openaimodel.py:350-413.  qkv [B, 3C, T] -> [B, C, T]."""
    bs, width, length = qkv.shape
    ch = width // (3 * heads)
    if not new_order:
        q, k, v = qkv.reshape(bs * heads, ch * 3, length).split(ch, dim=1)
    else:
        q, k, v = (z.reshape(bs * heads, ch, length) for z in qkv.chunk(3, dim=1))
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w, dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v)
    return a.reshape(bs, -1, length)


def _attnblock(sd, p, x, d, new_order):
    # openaimodel.py:321-327
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(xf, sd, p + ".norm"), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    h = qkv_attention(qkv, d["heads"], new_order)
    h = F.conv1d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + h).reshape(b, c, hh, ww)


def _cross_attention(sd, p, x, context, heads):
    # CrossAttention.forward, base/modules/attention.py:166-192 (no mask): x [B,N,dim], context [B,M,cdim] or None
    ctx = x if context is None else context
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(ctx, sd[p + ".to_k.weight"])
    v = F.linear(ctx, sd[p + ".to_v.weight"])
    b, n, inner = q.shape
    d = inner // heads
    sp = lambda t: t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)
    q, k, v = sp(q), sp(k), sp(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    out = torch.einsum("bij,bjd->bid", sim.softmax(dim=-1), v)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, inner)
    return F.linear(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def _spatial_transformer(sd, p, x, context, d):
    # SpatialTransformer.forward + BasicTransformerBlock._forward + GEGLU feed-forward
    # (base/modules/attention.py:36-67, 196-264); GroupNorm eps 1e-6 (:79-80), LayerNorm eps 1e-5
    b, c, hh, ww = x.shape
    heads = d["heads"]
    ctx = None if context is None else context.flatten(2).transpose(1, 2)       # 'b c h w -> b (h w) c'
    h = F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    h = F.conv2d(h, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    h = h.flatten(2).transpose(1, 2)
    j = 0
    while f"{p}.transformer_blocks.{j}.norm1.weight" in sd:
        q = f"{p}.transformer_blocks.{j}"
        ln = lambda t, n: F.layer_norm(t, (t.shape[-1],), sd[f"{q}.{n}.weight"], sd[f"{q}.{n}.bias"], 1e-5)
        h = _cross_attention(sd, q + ".attn1", ln(h, "norm1"), None, heads) + h
        h = _cross_attention(sd, q + ".attn2", ln(h, "norm2"), ctx, heads) + h
        u = F.linear(ln(h, "norm3"), sd[q + ".ff.net.0.proj.weight"], sd[q + ".ff.net.0.proj.bias"])
        a, gate = u.chunk(2, dim=-1)
        h = F.linear(a * F.gelu(gate), sd[q + ".ff.net.2.weight"], sd[q + ".ff.net.2.bias"]) + h
        j += 1
    h = h.transpose(1, 2).reshape(b, -1, hh, ww)
    return F.conv2d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"]) + x


def _run_layers(sd, prefix, layers, h, emb, cfg, context=None):
    for j, (kind, d) in enumerate(layers):
        p = f"{prefix}.{j}"
        if kind == "attn" and getattr(cfg, "use_spatial_transformer", False):
            h = _spatial_transformer(sd, p, h, context, d)
            continue
        if kind == "conv":
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "res":
            h = _resblock(sd, p, h, emb, d, cfg.use_scale_shift_norm)
        elif kind == "attn":
            h = _attnblock(sd, p, h, d, cfg.use_new_attention_order)
        elif kind == "downsample":
            if d["use_conv"]:
                h = F.conv2d(h, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
            else:
                h = F.avg_pool2d(h, 2)
        elif kind == "upsample":
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            if d["use_conv"]:
                h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
        else:
            raise ValueError(kind)
    return h


def unet_forward(sd, cfg, x, t, context=None, prefix=""):
    """UNetModel.forward (openaimodel.py:721-759) on a state dict ``sd``.

    ``prefix`` is e.g. "denoise_fn." when ``sd`` is a BrownianBridgeModel state dict.
    """
    if prefix:
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    dt = x.dtype
    plan = _block_plan(cfg)
    t_emb = timestep_embedding(t, cfg.model_channels).to(dt)
    emb = F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    if cfg.condition_key != "nocond":
        x = torch.cat([x, context], dim=1)
    h = x
    hs = []
    for i, layers in enumerate(plan["input"]):
        h = _run_layers(sd, f"input_blocks.{i}", layers, h, emb, cfg, context)
        hs.append(h)
    h = _run_layers(sd, "middle_block", plan["middle"], h, emb, cfg, context)
    for i, layers in enumerate(plan["output"]):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(sd, f"output_blocks.{i}", layers, h, emb, cfg, context)
    h = F.silu(_gn(h, sd, "out.0"))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


# --------------------------------------------------------------------------------------
# Whole-path drivers
# --------------------------------------------------------------------------------------


def p_sample(sd, cfg, bufs, steps, i, x_t, y, context, noise, objective="grad", eta=1.0,
             clip_denoised=False, prefix="denoise_fn."):
    """One p_sample step with the noise supplied by the caller (BrownianBridgeModel.py:171-201)."""
    B = x_t.shape[0]
    t = torch.full((B,), int(steps[i]), dtype=torch.long)
    eps = unet_forward(sd, cfg, x_t, t, context, prefix)
    return p_sample_update(bufs, steps, i, x_t, y, eps, noise, objective, eta, clip_denoised)


def p_sample_loop(sd, cfg, bufs, steps, y, context=None, objective="grad", eta=1.0,
                  clip_denoised=True, generator=None, prefix="denoise_fn.", n_steps=None):
    """p_sample_loop (BrownianBridgeModel.py:203-221); noise drawn with torch.randn like the
    reference's randn_like so a shared global seed reproduces it."""
    if cfg.condition_key == "nocond":
        context = None
    else:
        context = y if context is None else context
    img = y
    n = len(steps) if n_steps is None else n_steps
    for i in range(n):
        if int(steps[i]) == 0:
            noise = None
        else:
            noise = torch.randn(img.shape, dtype=torch.float32, generator=generator).to(img.dtype)
        img, _ = p_sample(sd, cfg, bufs, steps, i, img, y, context,
                          noise if noise is not None else torch.zeros_like(img),
                          objective, eta, clip_denoised, prefix)
    return img


def p_losses(sd, cfg, bufs, x0, y, context, t, noise, objective="grad", loss_type="l1",
             prefix="denoise_fn."):
    """p_losses (BrownianBridgeModel.py:98-126) with t and noise supplied."""
    x_t, obj = q_sample(bufs, x0, y, t, noise, objective)
    rec = unet_forward(sd, cfg, x_t, t, context, prefix)
    if loss_type == "l1":
        loss = (obj - rec).abs().mean()
    elif loss_type == "l2":
        loss = F.mse_loss(obj, rec)
    else:
        raise NotImplementedError
    return loss, predict_x0(bufs, x_t, y, t, rec, objective)


# --------------------------------------------------------------------------------------
# Per-kernel restatements in the product's device layout (NHWC, split-bf16 operands).
# These define what each C-ABI entry point in include/bbdm_b200.h must compute.
# --------------------------------------------------------------------------------------


def bf16_split(x):
    """x (fp32) -> (hi, lo) with hi = bf16(x), lo = bf16(x - hi), both returned as fp32."""
    hi = x.to(torch.bfloat16).to(torch.float32)
    lo = (x - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def op_gn_stats(x_nhwc, groups=32, eps=1e-5):
    """x [B,H,W,C] -> mean [B,G], rstd [B,G] (biased variance, like nn.GroupNorm)."""
    B, H, W, C = x_nhwc.shape
    xg = x_nhwc.reshape(B, H * W, groups, C // groups).double()
    mean = xg.mean(dim=(1, 3))
    var = xg.var(dim=(1, 3), unbiased=False)
    return mean.float(), (1.0 / torch.sqrt(var + eps)).float()


def op_gn_act(x_nhwc, mean, rstd, gamma, beta, film_scale=None, film_shift=None, silu=True,
              resample=0):
    """GroupNorm-affine (+FiLM) (+SiLU) (+nearest-up 2x / 2x2 avg-pool) in NHWC, fp32 result.

    resample: 0 none, 1 up, 2 down -- applied AFTER the activation (openaimodel.py:259-264).
    """
    B, H, W, C = x_nhwc.shape
    G = mean.shape[1]
    m = mean.repeat_interleave(C // G, dim=1)[:, None, None, :]
    r = rstd.repeat_interleave(C // G, dim=1)[:, None, None, :]
    h = (x_nhwc - m) * r * gamma + beta
    if film_scale is not None:
        h = h * (1 + film_scale[:, None, None, :]) + film_shift[:, None, None, :]
    if silu:
        h = F.silu(h)
    return op_resample(h, resample)


def op_resample(x_nhwc, resample):
    if resample == 1:
        return x_nhwc.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    if resample == 2:
        B, H, W, C = x_nhwc.shape
        return x_nhwc.reshape(B, H // 2, 2, W // 2, 2, C).mean(dim=(2, 4))
    return x_nhwc


def op_conv_nhwc(a_nhwc, w_oihw, bias=None, residual=None):
    """Stride-1 'same' conv (3x3 or 1x1) on NHWC input with OIHW weight; exact in a's dtype."""
    pad = w_oihw.shape[-1] // 2
    o = F.conv2d(a_nhwc.permute(0, 3, 1, 2), w_oihw.to(a_nhwc.dtype),
                 None if bias is None else bias.to(a_nhwc.dtype), padding=pad)
    o = o.permute(0, 2, 3, 1)
    if residual is not None:
        o = o + residual
    return o.contiguous()


def op_conv_split3(a_nhwc, w_oihw, bias=None, residual=None):
    """What the split-bf16 tensor-core conv computes, evaluated in fp64:
    A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (the lo*lo term is dropped)."""
    ah, al = bf16_split(a_nhwc.float())
    wh, wl = bf16_split(w_oihw.float())
    d = torch.float64
    o = op_conv_nhwc(ah.to(d), wh.to(d)) + op_conv_nhwc(al.to(d), wh.to(d)) + \
        op_conv_nhwc(ah.to(d), wl.to(d))
    if bias is not None:
        o = o + bias.to(d)
    if residual is not None:
        o = o + residual.to(d)
    return o


def op_attention_nhwc(qkv_btc, heads, new_order=False):
    """qkv [B,T,3C] (channel order as produced by the qkv 1x1 conv) -> [B,T,C]."""
    return qkv_attention(qkv_btc.permute(0, 2, 1), heads, new_order).permute(0, 2, 1).contiguous()


# ---------------------------------------------------------------------------------------------------
# VQGAN ends of the latent models: functional restatement on a state_dict.
#   Encoder / Decoder  model/VQGAN/model.py:342-537 ; ResnetBlock :76-138 ; AttnBlock :140-192 ;
#   Downsample :55-73 ; Upsample :38-53 ; VectorQuantizer2.forward  quantize.py:271-312 ;
#   VQModel.decode  vqgan.py:75-78.
# ---------------------------------------------------------------------------------------------------
def _vq_gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _vq_conv(sd, p, x, stride=1, padding=None):
    w = sd[p + ".weight"]
    return F.conv2d(x, w, sd[p + ".bias"], stride=stride, padding=w.shape[2] // 2 if padding is None else padding)


def _vq_resnet(sd, p, x):
    h = _vq_conv(sd, p + ".conv1", F.silu(_vq_gn(sd, p + ".norm1", x)))
    h = _vq_conv(sd, p + ".conv2", F.silu(_vq_gn(sd, p + ".norm2", h)))
    if p + ".nin_shortcut.weight" in sd:
        x = _vq_conv(sd, p + ".nin_shortcut", x)
    elif p + ".conv_shortcut.weight" in sd:
        x = _vq_conv(sd, p + ".conv_shortcut", x)
    return x + h


def _vq_attn(sd, p, x):
    b, c, hh, ww = x.shape
    h = _vq_gn(sd, p + ".norm", x)
    q = _vq_conv(sd, p + ".q", h).reshape(b, c, -1)
    k = _vq_conv(sd, p + ".k", h).reshape(b, c, -1)
    v = _vq_conv(sd, p + ".v", h).reshape(b, c, -1)
    w = torch.softmax(torch.einsum("bct,bcs->bts", q, k) * (int(c) ** (-0.5)), dim=2)     # [b, query, key]
    o = torch.einsum("bcs,bts->bct", v, w).reshape(b, c, hh, ww)
    return x + _vq_conv(sd, p + ".proj_out", o)


def _vq_mid(sd, p, h):
    h = _vq_resnet(sd, p + ".block_1", h)
    h = _vq_attn(sd, p + ".attn_1", h)
    return _vq_resnet(sd, p + ".block_2", h)


def vqgan_encode(sd, dd, x, quant_conv=True):
    """encoder(x) [-> quant_conv]; dd = the ddconfig dict."""
    n_res, nrb = len(dd["ch_mult"]), dd["num_res_blocks"]
    h = _vq_conv(sd, "encoder.conv_in", x)
    for i in range(n_res):
        for j in range(nrb):
            h = _vq_resnet(sd, f"encoder.down.{i}.block.{j}", h)
            if f"encoder.down.{i}.attn.{j}.norm.weight" in sd:
                h = _vq_attn(sd, f"encoder.down.{i}.attn.{j}", h)
        if i != n_res - 1:
            h = _vq_conv(sd, f"encoder.down.{i}.downsample.conv", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = _vq_mid(sd, "encoder.mid", h)
    h = _vq_conv(sd, "encoder.conv_out", F.silu(_vq_gn(sd, "encoder.norm_out", h)))
    return _vq_conv(sd, "quant_conv", h) if quant_conv else h


def vqgan_quantize(sd, z):
    """(z_q NCHW, indices [B*H*W]): nearest codebook row, z + (e - z)."""
    cb = sd["quantize.embedding.weight"]
    zf = z.permute(0, 2, 3, 1).contiguous()
    flat = zf.view(-1, cb.shape[1])
    d = torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(cb ** 2, dim=1) - 2 * flat @ cb.t()
    idx = torch.argmin(d, dim=1)
    zq = cb[idx].view(zf.shape)
    zq = zf + (zq - zf)
    return zq.permute(0, 3, 1, 2).contiguous(), idx


def vqgan_decode(sd, dd, z, quant_conv_first=False):
    """[quant_conv ->] quantize -> post_quant_conv -> decoder."""
    n_res, nrb = len(dd["ch_mult"]), dd["num_res_blocks"]
    if quant_conv_first:
        z = _vq_conv(sd, "quant_conv", z)
    zq, idx = vqgan_quantize(sd, z)
    h = _vq_conv(sd, "decoder.conv_in", _vq_conv(sd, "post_quant_conv", zq))
    h = _vq_mid(sd, "decoder.mid", h)
    for i in reversed(range(n_res)):
        for j in range(nrb + 1):
            h = _vq_resnet(sd, f"decoder.up.{i}.block.{j}", h)
            if f"decoder.up.{i}.attn.{j}.norm.weight" in sd:
                h = _vq_attn(sd, f"decoder.up.{i}.attn.{j}", h)
        if i != 0:
            h = _vq_conv(sd, f"decoder.up.{i}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    return _vq_conv(sd, "decoder.conv_out", F.silu(_vq_gn(sd, "decoder.norm_out", h))), idx
