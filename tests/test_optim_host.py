"""Host logic of the multi-tensor Adam / EMA (bbdm_b200/optim.py) on CPU through the emulation backend: state_dict
compatibility with torch.optim.Adam, EMA interface of the reference class, overlay import path."""
import copy

import pytest
import torch
import torch.nn as nn

import conftest
from _emu_backend import EmuBackend
from bbdm_b200 import optim as O


@pytest.fixture(autouse=True)
def _emu(monkeypatch):
    monkeypatch.setattr(O.FusedAdam, "backend_factory", staticmethod(lambda: EmuBackend()))
    monkeypatch.setattr(O.FusedEMA, "backend_factory", staticmethod(lambda: EmuBackend()))


def _net(seed=0):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.GroupNorm(4, 8), nn.SiLU(), nn.Conv2d(8, 5, 1), nn.Flatten(),
                         nn.Linear(5 * 36, 7))


def test_fused_adam_matches_torch_adam_and_state_dict_roundtrip():
    a, b = _net(), _net()
    oa = torch.optim.Adam(a.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-4)
    ob = O.FusedAdam(b.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-4)
    x = torch.randn(4, 3, 6, 6)
    for it in range(6):
        for net, opt in ((a, oa), (b, ob)):
            opt.zero_grad(set_to_none=True)
            net(x).square().mean().backward()
            opt.step()
        if it == 2:   # checkpoint round trip in both directions (runners/BaseRunner.py:131-152)
            sa, sb = copy.deepcopy(oa.state_dict()), copy.deepcopy(ob.state_dict())
            assert sa["param_groups"][0].keys() == sb["param_groups"][0].keys()
            assert set(sa["state"][0]) == set(sb["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
            oa.load_state_dict(sb)
            ob.load_state_dict(sa)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-8)
    assert float(ob.state_dict()["state"][0]["step"]) == 6.0


def test_fused_ema_interface_matches_reference_semantics():
    net = _net()
    ema = O.FusedEMA(0.9)
    ema.register(net)
    names = [n for n, p in net.named_parameters() if p.requires_grad]
    assert list(ema.shadow) == names
    ref = {n: p.data.clone() for n, p in net.named_parameters()}
    with torch.no_grad():
        for p in net.parameters():
            p.add_(1.0)
    ema.update(net, with_decay=True)
    for n, p in net.named_parameters():
        assert torch.equal(ema.shadow[n], (1.0 - 0.9) * p.data + 0.9 * ref[n])     # the reference's expression
    ema.update(net, with_decay=False)
    for n, p in net.named_parameters():
        assert torch.equal(ema.shadow[n], p.data)
    before = {n: p.data.clone() for n, p in net.named_parameters()}
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(2.0)
    ema.apply_shadow(net)
    for n, p in net.named_parameters():
        assert torch.equal(p.data, before[n])
    ema.restore(net)
    for n, p in net.named_parameters():
        assert torch.equal(p.data, 2.0 * before[n])
    # checkpoint path: the runner assigns a loaded dict, then reset_device (BaseRunner.py:125-126)
    ema.shadow = {n: torch.full_like(p.data, 3.0) for n, p in net.named_parameters()}
    ema.reset_device(net)
    ema.update(net, with_decay=True)
    for n, p in net.named_parameters():
        assert torch.allclose(ema.shadow[n], 0.1 * p.data + 0.9 * 3.0)


def test_adam_step_with_fused_ema_update():
    net = _net()
    ema = O.FusedEMA(0.99)
    ema.register(net)
    opt = O.FusedAdam(net.parameters(), lr=1e-2)
    s0 = {n: v.clone() for n, v in ema.shadow.items()}
    net(torch.randn(2, 3, 6, 6)).square().mean().backward()
    opt.step(ema=ema, ema_update=True)
    for n, p in net.named_parameters():
        assert torch.allclose(ema.shadow[n], 0.01 * p.data + 0.99 * s0[n], rtol=1e-6, atol=1e-8)


def test_overlay_module_exports_the_fused_ema():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("overlay_ema", os.path.join(conftest.REPO, "runners", "base", "EMA.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.EMA is O.FusedEMA
    assert not os.path.exists(os.path.join(conftest.REPO, "runners", "__init__.py"))
    assert not os.path.exists(os.path.join(conftest.REPO, "runners", "base", "__init__.py"))


def test_cpu_parameters_fail_loudly_without_emulation(monkeypatch):
    from bbdm_b200 import cabi
    monkeypatch.setattr(O.FusedEMA, "backend_factory", staticmethod(lambda: type("B", (), {"requires_cuda": True})()))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        O.FusedEMA(0.9).register(_net())
