"""Host logic of the cond-stage SpatialRescaler (bbdm_b200/cond.py; reference
model/BrownianBridge/base/modules/encoders/modules.py:106-134): the no-grad path is one backend call whose arithmetic
(2x2 expression of the bilinear kernel at scale 0.5, then the 1x1 map) equals the reference's op chain; with autograd
enabled the stock ops stay the graph."""
import pytest
import torch
import torch.nn.functional as F

from bbdm_b200.cond import SpatialRescaler
from _emu_backend import EmuBackend


@pytest.fixture
def emu(monkeypatch):
    be = EmuBackend()
    monkeypatch.setattr(SpatialRescaler, "backend_factory", staticmethod(lambda: be))
    return be


def _reference_chain(m, x):
    for _ in range(m.n_stages):
        x = F.interpolate(x, scale_factor=m.multiplier, mode=m.method)
    return m.channel_mapper(x) if m.remap_output else x


@pytest.mark.parametrize("n_stages,out_channels,bias,shape", [
    (2, 3, False, (2, 3, 64, 64)),          # Template-LBBDM CondStageParams
    (1, None, False, (1, 3, 37, 51)),       # odd sizes: floor per stage
    (3, 8, True, (2, 5, 40, 72)),
    (0, 4, True, (1, 3, 9, 7)),
])
def test_native_path_matches_reference_chain(emu, n_stages, out_channels, bias, shape):
    torch.manual_seed(n_stages)
    m = SpatialRescaler(n_stages=n_stages, in_channels=shape[1], out_channels=out_channels, bias=bias).eval()
    x = torch.randn(shape)
    with torch.no_grad():
        got = m(x)
        want = _reference_chain(m, x)
    assert emu.calls == ["spatial_rescale"]
    assert got.shape == want.shape
    assert (got - want).abs().max() <= 1e-6 * max(1.0, float(want.abs().max()))


def test_autograd_path_is_stock_ops(emu):
    m = SpatialRescaler(n_stages=2, in_channels=3, out_channels=3)
    x = torch.randn(1, 3, 16, 16)
    m(x).sum().backward()                                   # channel_mapper trains with the UNet
    assert emu.calls == [] and m.channel_mapper.weight.grad is not None


def test_unsupported_variant_warns_and_uses_stock_ops(emu):
    SpatialRescaler._warned = False
    m = SpatialRescaler(n_stages=1, method="nearest", in_channels=3)
    with torch.no_grad(), pytest.warns(UserWarning, match="stock PyTorch"):
        y = m(torch.randn(1, 3, 8, 8))
    assert y.shape == (1, 3, 4, 4) and emu.calls == []


def test_cpu_tensors_never_reach_the_library():
    m = SpatialRescaler(n_stages=1, in_channels=3)
    with torch.no_grad():
        assert m(torch.randn(1, 3, 8, 8)).shape == (1, 3, 4, 4)
    assert m._be is None
