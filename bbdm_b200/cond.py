"""Conditioning-stage module of the latent model.

``SpatialRescaler``: n_stages x interpolate(scale 0.5) + optional bias-free 1x1 channel map --
the only cond-stage encoder reachable from LatentBrownianBridgeModel (reference
model/BrownianBridge/base/modules/encoders/modules.py:106-134).  Tiny; plain PyTorch.
The parameter name ``channel_mapper`` matches the reference so checkpoints load.
"""
import torch.nn as nn
import torch.nn.functional as F

_METHODS = ('nearest', 'linear', 'bilinear', 'trilinear', 'bicubic', 'area')


class SpatialRescaler(nn.Module):
    def __init__(self, n_stages=1, method='bilinear', multiplier=0.5, in_channels=3, out_channels=None,
                 bias=False):
        super().__init__()
        assert n_stages >= 0 and method in _METHODS
        self.n_stages, self.method, self.multiplier = n_stages, method, multiplier
        self.remap_output = out_channels is not None
        if self.remap_output:
            self.channel_mapper = nn.Conv2d(in_channels, out_channels, 1, bias=bias)

    def forward(self, x):
        for _ in range(self.n_stages):
            x = F.interpolate(x, scale_factor=self.multiplier, mode=self.method)
        return self.channel_mapper(x) if self.remap_output else x

    def encode(self, x):
        return self(x)
