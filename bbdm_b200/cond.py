"""Conditioning-stage module of the latent model.

``SpatialRescaler``: n_stages x interpolate(scale 0.5) + optional bias-free 1x1 channel map --
the only cond-stage encoder reachable from LatentBrownianBridgeModel (reference
model/BrownianBridge/base/modules/encoders/modules.py:106-134).  The parameter name ``channel_mapper`` matches the
reference so checkpoints load.

No-grad CUDA calls (``sample`` / ``sample_to_eval``: the context is built once per batch and concatenated by the UNet
stem at every step) run as ONE kernel, ``bbdm_spatial_rescale``: the bilinear kernel at scale 0.5 has both weights
exactly 0.5, so the n stages and the channel map collapse into one pass that writes the NCHW context directly.
With autograd enabled (training: ``channel_mapper`` is optimised together with the UNet,
LatentBrownianBridgeModel.py:42-49) the stock PyTorch ops below are the graph.
"""
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import cabi

_METHODS = ('nearest', 'linear', 'bilinear', 'trilinear', 'bicubic', 'area')
NATIVE_MAX_STAGES, NATIVE_MAX_CHANNELS = 4, 16


def _product_backend():
    return cabi.CudaBackend()


class SpatialRescaler(nn.Module):
    backend_factory = staticmethod(_product_backend)      # tests substitute the emulation backend (CPU tensors)
    _warned = False

    def __init__(self, n_stages=1, method='bilinear', multiplier=0.5, in_channels=3, out_channels=None,
                 bias=False):
        super().__init__()
        assert n_stages >= 0 and method in _METHODS
        self.n_stages, self.method, self.multiplier = n_stages, method, multiplier
        self.remap_output = out_channels is not None
        if self.remap_output:
            self.channel_mapper = nn.Conv2d(in_channels, out_channels, 1, bias=bias)
        self._be = None

    def _native_ok(self, x):
        return (x.dim() == 4 and x.dtype == torch.float32 and self.method == 'bilinear' and self.multiplier == 0.5
                and self.n_stages <= NATIVE_MAX_STAGES and x.shape[1] <= NATIVE_MAX_CHANNELS
                and min(x.shape[2], x.shape[3]) >> self.n_stages > 0)

    def _native(self, x):
        if self._be is None:
            self._be = self.backend_factory()
        x = x.detach().contiguous()
        B, C, H, W = x.shape
        wt = bs = None
        if self.remap_output:
            cm = self.channel_mapper
            wt = cm.weight.detach().reshape(cm.out_channels, C).contiguous()
            bs = None if cm.bias is None else cm.bias.detach().contiguous()
        out = torch.empty((B, C if wt is None else wt.shape[0], H >> self.n_stages, W >> self.n_stages),
                          dtype=torch.float32, device=x.device)
        self._be.spatial_rescale(x, self.n_stages, wt, bs, out)
        return out

    def forward(self, x):
        if (x.is_cuda or type(self).backend_factory is not _product_backend) and not torch.is_grad_enabled():
            if self._native_ok(x):
                return self._native(x)
            if not SpatialRescaler._warned:
                SpatialRescaler._warned = True
                warnings.warn(f"bbdm_b200: SpatialRescaler(method={self.method!r}, multiplier={self.multiplier}, "
                              f"n_stages={self.n_stages}) runs on stock PyTorch kernels (the native kernel covers "
                              f"bilinear x0.5, <= {NATIVE_MAX_STAGES} stages, <= {NATIVE_MAX_CHANNELS} channels)",
                              stacklevel=2)
        for _ in range(self.n_stages):
            x = F.interpolate(x, scale_factor=self.multiplier, mode=self.method)
        return self.channel_mapper(x) if self.remap_output else x

    def encode(self, x):
        return self(x)
