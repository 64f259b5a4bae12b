"""-m gpu: the Winograd data-gradient route at REAL gradient magnitudes.

Its operand planes are fp16 pairs; a loss gradient (1e-4 ... 1e-8 after the mean over the batch) sits below fp16's
normal range, so ``train._conv_backward`` normalises dY by a power of two before the transform and scales the result
back.  The unit-scale dY of the other gradient tests cannot see that step; these two do (found with the LBBDM-f4 UNet
under the CPU emulation: 4e-3 per layer and 2.5e-2 on the earliest layers without the normalisation).
(File name: sorts after every other GPU test file.)"""
import pytest
import torch
import torch.nn.functional as F

from _recipe import UNET_CONFIGS, bb_namespace, fill_state_dict, rel_dev

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (scale * torch.randn(shape, generator=g)).float()


@pytest.mark.parametrize("gscale", [3e-7, 2e-3])
def test_winograd_data_gradient_at_loss_gradient_magnitudes(gscale):
    from bbdm_b200 import train
    from bbdm_b200.train import GNActConv2dFn
    B, H, W, C, Cout = 8, 32, 32, 256, 256                         # 512 tiles: Winograd forward + data gradient
    mk = lambda t: t.to(DEV).requires_grad_(True)
    x = mk(rnd((B, C, H, W), 10) + 0.2)
    gamma, beta = mk(1 + 0.1 * rnd((C,), 11)), mk(0.1 * rnd((C,), 12))
    scale, shift = mk(0.3 * rnd((B, C), 13)), mk(0.3 * rnd((B, C), 14))
    w, b = mk(rnd((Cout, C, 3, 3), 15, 0.05)), mk(rnd((Cout,), 16, 0.1))
    gy = rnd((B, Cout, H, W), 17, gscale).to(DEV)
    assert train._wino_ok(train.backend(), B, H, W, C, Cout, 3)
    y = GNActConv2dFn.apply(x, gamma, beta, scale, shift, w, b, 0, None)
    y.backward(gy)
    train.backend().check_fault()
    d = lambda t: t.detach().double().cpu().requires_grad_(True)
    xd, gd, bd, sd, hd, wd, bbd = d(x), d(gamma), d(beta), d(scale), d(shift), d(w), d(b)
    h = F.group_norm(xd, 32, gd, bd, 1e-5) * (1 + sd[:, :, None, None]) + hd[:, :, None, None]
    F.conv2d(F.silu(h), wd, bbd, padding=1).backward(gy.double().cpu())
    devs = {n: rel_dev(a.grad, r.grad) for n, a, r in [("x", x, xd), ("gamma", gamma, gd), ("beta", beta, bd), ("w", w, wd),
                                                       ("b", b, bbd), ("scale", scale, sd), ("shift", shift, hd)]}
    print(f"\n[winograd dgrad, |dY| ~ {gscale:g}] " + " ".join(f"{k} {v:.2e}" for k, v in devs.items()))
    assert max(devs.values()) < 1e-4, devs


def test_lbbdm_f4_training_step_with_winograd_route_matches_library_graph():
    """One training step of the LBBDM-f4 UNet (BASELINE configs[2] architecture, 2 samples): with the tile threshold
    lowered its 512-channel 32x32 level trains on the Winograd route, as it does at the benchmark batch of 32.
    Gradients of every parameter vs the stock fp32 graph of the same modules."""
    import bbdm_b200.unet as U
    from bbdm_b200 import train
    from model.BrownianBridge.BrownianBridgeModel import BrownianBridgeModel
    net = BrownianBridgeModel(bb_namespace(UNET_CONFIGS["lbbdm_f4"])).train()
    shapes = {k: tuple(v.shape) for k, v in net.denoise_fn.state_dict().items()}
    net.denoise_fn.load_state_dict(fill_state_dict(shapes, seed=1234))
    net = net.to(DEV)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 64, 64, generator=g).clamp_(-1, 1).to(DEV)
    y = torch.randn(2, 3, 64, 64, generator=g).clamp_(-1, 1).to(DEV)
    nz = torch.randn(2, 3, 64, 64, generator=g).to(DEV)
    t = torch.tensor([100, 700], device=DEV)
    old_tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    old_tiles = train.WINO_MIN_TILES
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    train.WINO_MIN_TILES = 128
    try:
        assert train._wino_ok(train.backend(), 2, 32, 32, 512, 512, 3)
        res = {}
        for native in (True, False):
            U.NATIVE_TRAIN_CONV = native
            net.zero_grad(set_to_none=True)
            loss, _ = net.p_losses(x, y, y, t, nz)
            loss.backward()
            res[native] = (float(loss), {n: p.grad.detach().clone() for n, p in net.denoise_fn.named_parameters()})
        train.backend().check_fault()
    finally:
        U.NATIVE_TRAIN_CONV = True
        train.WINO_MIN_TILES = old_tiles
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_tf32
    assert abs(res[True][0] - res[False][0]) < 1e-4 * abs(res[False][0])
    devs = {n: rel_dev(res[True][1][n], res[False][1][n]) for n in res[False][1]}
    wname = max(devs, key=devs.get)
    print(f"\n[train lbbdm_f4, Winograd forced] loss {res[True][0]:.6f} / {res[False][0]:.6f}; worst grad {wname} {devs[wname]:.3e}")
    assert devs[wname] < 1e-3, (wname, devs[wname])
