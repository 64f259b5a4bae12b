"""SpatialTransformer / cross-attention conditioning -- parameter tree.

Same parameter names and shapes as the reference modules
(/root/reference/model/BrownianBridge/base/modules/attention.py:36-264: GEGLU, FeedForward, CrossAttention,
BasicTransformerBlock, SpatialTransformer), so checkpoints, EMA and ``weights_init`` (which keys on the class
names ``Linear`` / ``Conv2d``) work unchanged.  The ``forward`` methods below are the training / autograd graph in
stock PyTorch ops; no-grad CUDA calls are executed by ``bbdm_b200.engine.UNetEngine._spatial_transformer`` on the
sm_100a kernels (GroupNorm + 1x1 projections and every Linear on the tcgen05 GEMM, LayerNorm / GEGLU as fused
operand-producing passes, self- and cross-attention on the flash kernels).

Reference semantics kept: the UNet passes the SAME 4-D ``context`` tensor it concatenates to the input
(openaimodel.py:741-748) to every transformer, where it is flattened to ``b (h w) c`` (attention.py:171-172), so
``context_dim`` is its channel count and the cross-attention runs over all of its pixels.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _zero(m):
    for p in m.parameters():
        p.detach().zero_()
    return m


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.0):
        super().__init__()
        inner = int(dim * mult)
        dim_out = dim if dim_out is None else dim_out
        project_in = GEGLU(dim, inner) if glu else nn.Sequential(nn.Linear(dim, inner), nn.GELU())
        self.glu = glu
        # indices 0 and 2 carry the parameters, like the reference's Sequential (attention.py:62-66)
        self.net = nn.Sequential(project_in, nn.Dropout(dropout), nn.Linear(inner, dim_out))

    def forward(self, x):
        return self.net(x)


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        context_dim = query_dim if context_dim is None else context_dim
        self.scale = dim_head ** -0.5
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))

    def forward(self, x, context=None):
        h = self.heads
        if context is not None:
            context = context.flatten(2).transpose(1, 2)          # 'b c h w -> b (h w) c'
        else:
            context = x
        q, k, v = self.to_q(x), self.to_k(context), self.to_v(context)
        b, n, _ = q.shape
        split = lambda t: t.reshape(b, t.shape[1], h, -1).permute(0, 2, 1, 3).reshape(b * h, t.shape[1], -1)
        q, k, v = split(q), split(k), split(v)
        attn = (torch.einsum("bid,bjd->bij", q, k) * self.scale).softmax(dim=-1)
        out = torch.einsum("bij,bjd->bid", attn, v)
        out = out.reshape(b, h, n, -1).permute(0, 2, 1, 3).reshape(b, n, -1)
        return self.to_out(out)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True):
        super().__init__()
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)      # self-attention
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                                    dropout=dropout)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)

    def forward(self, x, context=None):
        # (the reference wraps this in its checkpoint(): same values)
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context=context) + x
        return self.ff(self.norm3(x)) + x


class SpatialTransformer(nn.Module):
    """GroupNorm(eps 1e-6) -> 1x1 proj_in -> depth x BasicTransformerBlock over 'b (h w) c' -> 1x1 proj_out -> + x
    (attention.py:218-264)."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None):
        super().__init__()
        self.in_channels, self.n_heads, self.d_head, self.context_dim = in_channels, n_heads, d_head, context_dim
        inner = n_heads * d_head
        self.norm = nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, kernel_size=1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, n_heads, d_head, dropout=dropout, context_dim=context_dim) for _ in range(depth)])
        self.proj_out = _zero(nn.Conv2d(inner, in_channels, kernel_size=1))

    _warned = False

    def forward(self, x, context=None):
        if x.is_cuda and not SpatialTransformer._warned:
            SpatialTransformer._warned = True
            import warnings
            warnings.warn("bbdm_b200: SpatialTransformer blocks train on stock PyTorch kernels (their inference path is "
                          "native: UNetEngine._spatial_transformer)", stacklevel=2)
        b, c, h, w = x.shape
        x_in = x
        x = self.proj_in(self.norm(x))
        x = x.flatten(2).transpose(1, 2)
        for blk in self.transformer_blocks:
            x = blk(x, context=context)
        x = x.transpose(1, 2).reshape(b, -1, h, w)
        return self.proj_out(x) + x_in
