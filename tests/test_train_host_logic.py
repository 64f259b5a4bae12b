"""Host logic of the training path (bbdm_b200/train.py autograd Functions: operand plumbing, GroupNorm / FiLM
gradient algebra, resampling adjoints, residual pass-through, attention core) on CPU, with the oracle-backed
backend emulation standing in for the kernels (tests only) -- against fp64 PyTorch autograd of the same expression."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _emu_backend import EmuBackend
from _recipe import UNET_CONFIGS, bb_namespace, fill_state_dict, rel_dev


@pytest.fixture()
def emu():
    from bbdm_b200 import train
    be = EmuBackend()
    train.set_backend(be)
    yield be
    train.set_backend(None)


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (scale * torch.randn(shape, generator=g)).float()


def d64(t):
    return None if t is None else t.detach().double().requires_grad_(True)


@pytest.mark.parametrize("k,bias", [(3, True), (1, False)])
def test_conv2d_function(emu, k, bias):
    from bbdm_b200.train import Conv2dFn
    x = rnd((2, 64, 8, 8), 1).requires_grad_(True)
    w = rnd((128, 64, k, k), 2, 0.05).requires_grad_(True)
    b = rnd((128,), 3, 0.1).requires_grad_(True) if bias else None
    gy = rnd((2, 128, 8, 8), 4, 0.2)
    y = Conv2dFn.apply(x, w, b)
    y.backward(gy)
    xd, wd, bd = d64(x), d64(w), d64(b)
    yd = F.conv2d(xd, wd, bd, padding=k // 2)
    yd.backward(gy.double())
    assert rel_dev(y, yd) < 3e-5
    for a, r in ((x, xd), (w, wd)) + (((b, bd),) if bias else ()):
        assert rel_dev(a.grad, r.grad) < 5e-5
    assert {"pack_weight_split_both", "conv_umma", "split_grad", "conv_wgrad"} <= set(emu.calls)


@pytest.mark.parametrize("film,gscale", [(True, 0.2), (False, 0.2), (True, 3e-7)])
def test_gn_act_conv_function_winograd_route(emu, film, gscale, monkeypatch):
    """Forward and data gradient on the Winograd path (input transform that also emits the activated split planes for
    the weight gradient, 36 position GEMMs, output transform with bias + residual; dY transformed in identity mode with
    the flipped / channel-swapped kernel), forced on for a 64-channel 32x32 case (128 tiles).  gscale 3e-7 is the
    magnitude of real loss gradients: below fp16's normal range unless dY is normalised before the transform."""
    from bbdm_b200 import train
    from bbdm_b200.train import GNActConv2dFn
    monkeypatch.setattr(train, "WINO_MIN_C", 64)
    monkeypatch.setattr(train, "WINO_MIN_TILES", 128)
    B, C, H, W, Cout = 2, 64, 32, 32, 128
    x = (rnd((B, C, H, W), 10) + 0.2).requires_grad_(True)
    gamma, beta = (1 + 0.1 * rnd((C,), 11)).requires_grad_(True), (0.1 * rnd((C,), 12)).requires_grad_(True)
    scale = (0.3 * rnd((B, C), 13)).requires_grad_(True) if film else None
    shift = (0.3 * rnd((B, C), 14)).requires_grad_(True) if film else None
    w, b = rnd((Cout, C, 3, 3), 15, 0.05).requires_grad_(True), rnd((Cout,), 16, 0.1).requires_grad_(True)
    res = rnd((B, Cout, H, W), 18).requires_grad_(True) if film else None
    gy = rnd((B, Cout, H, W), 17, gscale)
    y = GNActConv2dFn.apply(x, gamma, beta, scale, shift, w, b, 0, res, True)
    y.backward(gy)
    assert emu.calls.count("wino_input") == 2 and emu.calls.count("wino_output") == 2 and "conv_wgrad" in emu.calls
    xd, gd, bd, sd, hd, wd, bbd, rd = (d64(t) for t in (x, gamma, beta, scale, shift, w, b, res))
    h = F.group_norm(xd, 32, gd, bd, 1e-5)
    if film:
        h = h * (1 + sd[:, :, None, None]) + hd[:, :, None, None]
    yd = F.conv2d(F.silu(h), wd, bbd, padding=1)
    if res is not None:
        yd = yd + rd
    yd.backward(gy.double())
    assert rel_dev(y, yd) < 3e-5
    for name, a, r in [("x", x, xd), ("gamma", gamma, gd), ("beta", beta, bd), ("w", w, wd), ("b", b, bbd)]:
        assert rel_dev(a.grad, r.grad) < 1e-4, name
    if film:
        assert rel_dev(scale.grad, sd.grad) < 1e-4 and rel_dev(shift.grad, hd.grad) < 1e-4
        assert rel_dev(res.grad, rd.grad) < 1e-6


@pytest.mark.parametrize("film,resample,act", [(True, 0, True), (False, 1, True), (False, 2, True), (False, 0, False)])
def test_gn_act_conv_function(emu, film, resample, act):
    from bbdm_b200.train import GNActConv2dFn
    B, C, H, W, Cout = 2, 64, 8, 8, 64
    Ho, Wo = (2 * H, 2 * W) if resample == 1 else ((H // 2, W // 2) if resample == 2 else (H, W))
    x = (rnd((B, C, H, W), 10) + 0.2).requires_grad_(True)
    gamma, beta = (1 + 0.1 * rnd((C,), 11)).requires_grad_(True), (0.1 * rnd((C,), 12)).requires_grad_(True)
    scale = (0.3 * rnd((B, C), 13)).requires_grad_(True) if film else None
    shift = (0.3 * rnd((B, C), 14)).requires_grad_(True) if film else None
    w, b = rnd((Cout, C, 3, 3), 15, 0.05).requires_grad_(True), rnd((Cout,), 16, 0.1).requires_grad_(True)
    res = rnd((B, Cout, Ho, Wo), 18).requires_grad_(True) if film else None
    gy = rnd((B, Cout, Ho, Wo), 17, 0.2)
    y = GNActConv2dFn.apply(x, gamma, beta, scale, shift, w, b, resample, res, act)
    y.backward(gy)
    xd, gd, bd, sd, hd, wd, bbd, rd = (d64(t) for t in (x, gamma, beta, scale, shift, w, b, res))
    h = F.group_norm(xd, 32, gd, bd, 1e-5)
    if film:
        h = h * (1 + sd[:, :, None, None]) + hd[:, :, None, None]
    if act:
        h = F.silu(h)
    if resample == 1:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
    elif resample == 2:
        h = F.avg_pool2d(h, 2)
    yd = F.conv2d(h, wd, bbd, padding=1)
    if res is not None:
        yd = yd + rd
    yd.backward(gy.double())
    assert rel_dev(y, yd) < 3e-5
    pairs = [("x", x, xd), ("gamma", gamma, gd), ("beta", beta, bd), ("w", w, wd), ("b", b, bbd)]
    if film:
        pairs += [("scale", scale, sd), ("shift", shift, hd), ("residual", res, rd)]
    for name, a, r in pairs:
        assert rel_dev(a.grad, r.grad) < 1e-4, (name, rel_dev(a.grad, r.grad))


@pytest.mark.parametrize("C,heads,order", [(128, 2, 0), (64, 2, 1)])
def test_attention_core_function(emu, C, heads, order):
    from bbdm_b200.train import AttentionCoreFn
    B, H, W = 2, 4, 4
    qkv = rnd((B, 3 * C, H, W), 32, 1.2).requires_grad_(True)
    gy = rnd((B, C, H, W), 33, 0.3)
    y = AttentionCoreFn.apply(qkv, heads, order)
    y.backward(gy)
    qd = d64(qkv)
    T, D = H * W, C // heads
    q3 = qd.permute(0, 2, 3, 1).reshape(B, T, 3 * C)
    if order == 0:
        q, k, v = q3.view(B, T, heads, 3, D).unbind(3)
    else:
        q, k, v = q3.view(B, T, 3, heads, D).unbind(2)
    s = D ** -0.25
    att = torch.einsum("bthd,bshd->bhts", q * s, k * s).softmax(-1)
    od = torch.einsum("bhts,bshd->bthd", att, v).reshape(B, H, W, C).permute(0, 3, 1, 2)
    od.backward(gy.double())
    assert rel_dev(y, od) < 3e-5
    assert rel_dev(qkv.grad, qd.grad) < 5e-5
    assert ("attention_tc" if D == 64 else "attention") in emu.calls and "attention_bwd" in emu.calls


def test_small_conv_function(emu):
    from bbdm_b200.train import SmallConv2dFn
    x = rnd((2, 6, 8, 8), 40).requires_grad_(True)
    w, b = rnd((32, 6, 3, 3), 41, 0.1).requires_grad_(True), rnd((32,), 42, 0.1).requires_grad_(True)
    gy = rnd((2, 32, 8, 8), 43, 0.2)
    y = SmallConv2dFn.apply(x, w, b)
    y.backward(gy)
    xd, wd, bd = d64(x), d64(w), d64(b)
    yd = F.conv2d(xd, wd, bd, padding=1)
    yd.backward(gy.double())
    assert rel_dev(y, yd) < 1e-6
    for a, r in ((x, xd), (w, wd), (b, bd)):
        assert rel_dev(a.grad, r.grad) < 1e-5


@pytest.mark.parametrize("force_winograd", [False, True])
def test_training_step_host_logic_matches_torch_graph(emu, monkeypatch, force_winograd):
    """One full training step of the tensor-core-aligned small UNet: every native Function (conv, GN+act+conv with
    FiLM / resampling / fused skip, attention block) wired by unet.py vs the stock-PyTorch graph of the same modules.
    force_winograd: the 3x3 convs of the two upper levels take the Winograd route (forward and data gradient) with the
    model's REAL gradient magnitudes (1e-4 and below -- the case a unit-scale dY never exercises)."""
    import bbdm_b200.unet as U
    from bbdm_b200 import train
    from bbdm_b200.bridge import BridgeOps
    if force_winograd:
        monkeypatch.setattr(train, "WINO_MIN_C", 64)
        monkeypatch.setattr(train, "WINO_MIN_TILES", 16)
    from model.BrownianBridge.BrownianBridgeModel import BrownianBridgeModel
    monkeypatch.setattr(BridgeOps, "backend_factory", staticmethod(lambda: emu))          # q_sample
    g = {k: torch.from_numpy(v) if v.ndim else v for k, v in np.load(os.path.join(os.path.dirname(__file__), "golden", "mid_pixel.npz")).items()}
    net = BrownianBridgeModel(bb_namespace(UNET_CONFIGS["mid_pixel"])).train()
    shapes = {k: tuple(v.shape) for k, v in net.denoise_fn.state_dict().items()}
    net.denoise_fn.load_state_dict(fill_state_dict(shapes, seed=1234))
    x, y, t, nz = (g[k] for k in ("x", "y", "t", "q_noise"))
    res = {}
    for native in (True, False):
        U.NATIVE_TRAIN_CONV = native
        net.zero_grad(set_to_none=True)
        emu.calls.clear()
        loss, _ = net.p_losses(x, y, y, t, nz)
        loss.backward()
        res[native] = (float(loss.detach()), {n: p.grad.detach().clone() for n, p in net.denoise_fn.named_parameters()}, set(emu.calls))
    U.NATIVE_TRAIN_CONV = True
    assert {"conv_umma", "conv_wgrad", "gn_bwd_reduce", "gn_bwd_apply", "attention_bwd", "conv_wgrad_direct"} <= res[True][2]
    assert ("wino_input" in res[True][2]) == force_winograd
    assert not ({"conv_umma", "conv_wgrad"} & res[False][2])
    assert abs(res[True][0] - float(g["loss"])) < 2e-4 * abs(float(g["loss"]))          # reference-generated loss
    worst = max(rel_dev(res[True][1][n], res[False][1][n]) for n in res[False][1])
    assert worst < 3e-4, worst
    # gradients of the UNMODIFIED reference for the same step (tests/golden/make_golden.py --grads-only)
    gr = np.load(os.path.join(os.path.dirname(__file__), "golden", "mid_pixel_grads.npz"))
    assert abs(res[True][0] - float(gr["loss"])) < 2e-4 * abs(float(gr["loss"]))
    for k in gr.files:
        if k.startswith("grad:"):
            assert rel_dev(res[False][1][k[5:]], torch.from_numpy(gr[k])) < 2e-5, k      # stock graph of OUR modules == reference
            assert rel_dev(res[True][1][k[5:]], torch.from_numpy(gr[k])) < 3e-4, k       # native Functions (emulated kernels)
