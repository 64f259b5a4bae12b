"""-m gpu: the training path's native pieces -- gradient operand split, tcgen05 weight-gradient
GEMM, data gradient through the forward conv kernel, the autograd Function, and one full
training step against the stock-PyTorch graph."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _recipe import UNET_CONFIGS, bb_namespace, fill_state_dict, rel_dev
from oracle import bbdm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def be():
    from bbdm_b200 import cabi
    b = cabi.CudaBackend()
    yield b
    b.check_fault()


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (scale * torch.randn(shape, generator=g)).float()


@pytest.mark.parametrize("P,C", [(256, 64), (1000, 96), (4096, 128), (64, 200)])
def test_split_grad(be, P, C):
    x = rnd((P, C), 1)
    hi, lo = torch.empty((P, C), dtype=torch.bfloat16, device=DEV), torch.empty((P, C), dtype=torch.bfloat16, device=DEV)
    ht, lt = torch.empty((C, P), dtype=torch.bfloat16, device=DEV), torch.empty((C, P), dtype=torch.bfloat16, device=DEV)
    cs = torch.empty((C,), device=DEV)
    ws = torch.empty(((P + 63) // 64) * C, device=DEV)
    be.split_grad(x.to(DEV), hi, lo, ht, lt, cs, ws)
    h, l = O.bf16_split(x)
    assert torch.equal(hi.float().cpu(), h) and torch.equal(lo.float().cpu(), l)
    assert torch.equal(ht.float().cpu(), h.T.contiguous()) and torch.equal(lt.float().cpu(), l.T.contiguous())
    assert rel_dev(cs, x.double().sum(0)) < 1e-6


WG_CASES = [
    # B, H, W, Cin, Cout, k
    (2, 8, 8, 64, 64, 3),        # BN = 64 (single MN atom), 64-pixel box = one image
    (1, 16, 16, 128, 128, 3),    # BN = 128: two MN atoms (LBO)
    (2, 16, 16, 256, 128, 3),    # BN = 256: four MN atoms, 8 promotion warps
    (2, 64, 64, 64, 192, 3),     # box = one row of 64; Cout tile partially out of range
    (4, 4, 4, 128, 64, 3),       # box spans 4 images
    (2, 16, 16, 128, 256, 1),    # 1x1
    (1, 32, 32, 512, 512, 3),    # deeper K split
]


@pytest.mark.parametrize("case", WG_CASES)
def test_conv_wgrad(be, case):
    B, H, W, Cin, Cout, k = case
    a = rnd((B, H, W, Cin), 2)
    g = rnd((B, H, W, Cout), 3, 0.1)
    P = B * H * W
    a_hi, a_lo = (t.to(torch.bfloat16).to(DEV) for t in O.bf16_split(a))
    ht, lt = torch.empty((Cout, P), dtype=torch.bfloat16, device=DEV), torch.empty((Cout, P), dtype=torch.bfloat16, device=DEV)
    be.split_grad(g.reshape(P, Cout).to(DEV), None, None, ht, lt)
    _, fl = be.wgrad_workspace(B, H, W, Cin, Cout, k * k)
    ws = torch.empty(fl, device=DEV)
    dw = torch.full((Cout, Cin, k, k), float("nan"), device=DEV)
    be.conv_wgrad(ht, lt, a_hi, a_lo, B, H, W, Cin, Cout, k * k, dw, ws)
    torch.cuda.synchronize()
    be.check_fault()
    # exact gradient for the values the planes carry
    av = sum(O.bf16_split(a)).double().permute(0, 3, 1, 2).requires_grad_(False)
    gv = sum(O.bf16_split(g)).double().permute(0, 3, 1, 2)
    w = torch.zeros((Cout, Cin, k, k), dtype=torch.float64, requires_grad=True)
    F.conv2d(av, w, padding=k // 2).backward(gv)
    assert not torch.isnan(dw).any()
    assert rel_dev(dw, w.grad) < 2e-5, rel_dev(dw, w.grad)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,bias", [(2, 16, 16, 64, 128, 3, True), (1, 32, 32, 128, 64, 3, False),
                                                    (2, 8, 8, 256, 256, 1, True)])
def test_conv2d_function_gradients(B, H, W, Cin, Cout, k, bias):
    from bbdm_b200.train import Conv2dFn
    x = rnd((B, Cin, H, W), 4).to(DEV).requires_grad_(True)
    w = rnd((Cout, Cin, k, k), 5, 0.05).to(DEV).requires_grad_(True)
    b = rnd((Cout,), 6, 0.1).to(DEV).requires_grad_(True) if bias else None
    gy = rnd((B, Cout, H, W), 7, 0.2).to(DEV)
    y = Conv2dFn.apply(x, w, b)
    y.backward(gy)
    xd, wd = x.detach().double().cpu().requires_grad_(True), w.detach().double().cpu().requires_grad_(True)
    bd = None if b is None else b.detach().double().cpu().requires_grad_(True)
    yd = F.conv2d(xd, wd, bd, padding=k // 2)
    yd.backward(gy.double().cpu())
    assert rel_dev(y, yd) < 3e-5
    assert rel_dev(x.grad, xd.grad) < 3e-5
    assert rel_dev(w.grad, wd.grad) < 3e-5
    if bias:
        assert rel_dev(b.grad, bd.grad) < 1e-5


def test_training_step_native_convs_match_library_graph():
    """loss and parameter gradients of one training step: tensor-core conv path vs stock PyTorch."""
    import bbdm_b200.unet as U
    from model.BrownianBridge.BrownianBridgeModel import BrownianBridgeModel
    g = {k: torch.from_numpy(v) if v.ndim else v for k, v in np.load(os.path.join(os.path.dirname(__file__), "golden", "mid_pixel.npz")).items()}
    net = BrownianBridgeModel(bb_namespace(UNET_CONFIGS["mid_pixel"])).train()
    shapes = {k: tuple(v.shape) for k, v in net.denoise_fn.state_dict().items()}
    net.denoise_fn.load_state_dict(fill_state_dict(shapes, seed=1234))
    net = net.cuda()
    x, y, t, nz = (g[k].cuda() for k in ("x", "y", "t", "q_noise"))
    res = {}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for native in (True, False):
        U.NATIVE_TRAIN_CONV = native
        net.zero_grad(set_to_none=True)
        loss, _ = net.p_losses(x, y, y, t, nz)
        loss.backward()
        res[native] = (float(loss), {n: p.grad.detach().clone() for n, p in net.denoise_fn.named_parameters()})
    U.NATIVE_TRAIN_CONV = True
    assert abs(res[True][0] - float(g["loss"])) < 2e-4 * abs(float(g["loss"]))
    assert abs(res[True][0] - res[False][0]) < 1e-4 * abs(res[False][0])
    worst = max(rel_dev(res[True][1][n], res[False][1][n]) for n in res[False][1])
    print(f"\n[train] loss native {res[True][0]:.6f} library {res[False][0]:.6f}; worst grad rel dev {worst:.3e}")
    assert worst < 3e-4
    # ... and against gradients the UNMODIFIED reference computed on CPU (tests/golden/make_golden.py --grads-only)
    gr = np.load(os.path.join(os.path.dirname(__file__), "golden", "mid_pixel_grads.npz"))
    assert abs(res[True][0] - float(gr["loss"])) < 2e-4 * abs(float(gr["loss"]))
    devs = {k[5:]: rel_dev(res[True][1][k[5:]], torch.from_numpy(gr[k])) for k in gr.files if k.startswith("grad:")}
    wname = max(devs, key=devs.get)
    print(f"[train] vs reference gradient fixture: {len(devs)} tensors, worst {wname} {devs[wname]:.3e}")
    assert len(devs) >= 20 and devs[wname] < 3e-4


@pytest.mark.parametrize("B,H,W,C,Cout,film", [(2, 16, 16, 64, 128, True), (3, 8, 8, 128, 64, False), (2, 32, 32, 640, 128, True),
                                                (8, 32, 32, 256, 256, True),     # Winograd forward + data gradient (512 tiles)
                                                (2, 64, 64, 256, 512, False)])   # Winograd: 256 tiles per image
def test_gn_act_conv_function_gradients(B, H, W, C, Cout, film):
    """conv(silu(GN(x)*(1+scale)+shift)) fused Function: outputs and ALL gradients vs an fp64 torch graph."""
    from bbdm_b200 import cabi, train
    from bbdm_b200.train import GNActConv2dFn
    mk = lambda t: t.to(DEV).requires_grad_(True)
    x = mk(rnd((B, C, H, W), 10) + 0.2)
    gamma, beta = mk(1 + 0.1 * rnd((C,), 11)), mk(0.1 * rnd((C,), 12))
    scale = mk(0.3 * rnd((B, C), 13)) if film else None
    shift = mk(0.3 * rnd((B, C), 14)) if film else None
    w, b = mk(rnd((Cout, C, 3, 3), 15, 0.05)), mk(rnd((Cout,), 16, 0.1))
    gy = rnd((B, Cout, H, W), 17, 0.2).to(DEV)
    res = mk(rnd((B, Cout, H, W), 18)) if film else None            # fused "+ skip" operand
    assert train._wino_ok(train.backend(), B, H, W, C, Cout, 3) == (min(C, Cout) >= 256)
    y = GNActConv2dFn.apply(x, gamma, beta, scale, shift, w, b, 0, res)
    y.backward(gy)
    train.backend().check_fault()

    d = lambda t: None if t is None else t.detach().double().cpu().requires_grad_(True)
    xd, gd, bd, sd, hd, wd, bbd, rd = d(x), d(gamma), d(beta), d(scale), d(shift), d(w), d(b), d(res)
    h = F.group_norm(xd, 32, gd, bd, 1e-5)
    if film:
        h = h * (1 + sd[:, :, None, None]) + hd[:, :, None, None]
    yd = F.conv2d(F.silu(h), wd, bbd, padding=1)
    if res is not None:
        yd = yd + rd
    yd.backward(gy.double().cpu())
    assert rel_dev(y, yd) < 3e-5
    pairs = [("x", x, xd), ("gamma", gamma, gd), ("beta", beta, bd), ("w", w, wd), ("b", b, bbd)]
    if res is not None:
        pairs.append(("residual", res, rd))
    if film:
        pairs += [("scale", scale, sd), ("shift", shift, hd)]
    for name, a, r in pairs:
        assert rel_dev(a.grad, r.grad) < 5e-5, (name, rel_dev(a.grad, r.grad))


@pytest.mark.parametrize("resample", [1, 2])
def test_gn_act_conv_function_with_resampling(resample):
    """up / down ResBlock in_layers: GN -> SiLU -> (nearest-2x | 2x2 mean) -> conv, fused, all gradients."""
    from bbdm_b200.train import GNActConv2dFn
    B, C, Cout, Hs = 2, 128, 128, 16
    mk = lambda t: t.to(DEV).requires_grad_(True)
    x = mk(rnd((B, C, Hs, Hs), 20) + 0.1)
    gamma, beta = mk(1 + 0.1 * rnd((C,), 21)), mk(0.1 * rnd((C,), 22))
    w, b = mk(rnd((Cout, C, 3, 3), 23, 0.05)), mk(rnd((Cout,), 24, 0.1))
    H = Hs * 2 if resample == 1 else Hs // 2
    gy = rnd((B, Cout, H, H), 25, 0.2).to(DEV)
    y = GNActConv2dFn.apply(x, gamma, beta, None, None, w, b, resample)
    y.backward(gy)
    d = lambda t: t.detach().double().cpu().requires_grad_(True)
    xd, gd, bd, wd, bbd = d(x), d(gamma), d(beta), d(w), d(b)
    h = F.silu(F.group_norm(xd, 32, gd, bd, 1e-5))
    h = F.interpolate(h, scale_factor=2, mode="nearest") if resample == 1 else F.avg_pool2d(h, 2)
    yd = F.conv2d(h, wd, bbd, padding=1)
    yd.backward(gy.double().cpu())
    assert rel_dev(y, yd) < 3e-5
    for name, a, r in [("x", x, xd), ("gamma", gamma, gd), ("beta", beta, bd), ("w", w, wd), ("b", b, bbd)]:
        assert rel_dev(a.grad, r.grad) < 5e-5, (name, rel_dev(a.grad, r.grad))


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,need_dx", [(2, 16, 16, 6, 128, 3, False), (2, 16, 16, 128, 3, 3, True),
                                                       (1, 8, 12, 16, 16, 3, True), (2, 8, 8, 4, 64, 1, True)])
def test_small_conv_function_gradients(B, H, W, Cin, Cout, k, need_dx):
    """Stem / head style convolutions (few channels on one side): exact-fp32 forward, dgrad, wgrad."""
    from bbdm_b200.train import SmallConv2dFn
    x = rnd((B, Cin, H, W), 30).to(DEV).requires_grad_(need_dx)
    w = rnd((Cout, Cin, k, k), 31, 0.05).to(DEV).requires_grad_(True)
    b = rnd((Cout,), 32, 0.1).to(DEV).requires_grad_(True)
    gy = rnd((B, Cout, H, W), 33, 0.2).to(DEV)
    y = SmallConv2dFn.apply(x, w, b)
    y.backward(gy)
    xd = x.detach().double().cpu().requires_grad_(need_dx)
    wd, bd = w.detach().double().cpu().requires_grad_(True), b.detach().double().cpu().requires_grad_(True)
    yd = F.conv2d(xd, wd, bd, padding=k // 2)
    yd.backward(gy.double().cpu())
    assert rel_dev(y, yd) < 2e-6
    assert rel_dev(w.grad, wd.grad) < 2e-6 and rel_dev(b.grad, bd.grad) < 2e-6
    if need_dx:
        assert rel_dev(x.grad, xd.grad) < 2e-6


def _attention_ref(qkv, heads, order):
    """QKVAttentionLegacy / QKVAttention (openaimodel.py:350-413) on a [B,T,3C] tensor, any dtype."""
    B, T, C3 = qkv.shape
    Cc = C3 // 3
    D = Cc // heads
    if order == 0:
        q, k, v = qkv.view(B, T, heads, 3, D).unbind(3)                  # per-head q|k|v interleave
    else:
        q, k, v = qkv.view(B, T, 3, heads, D).unbind(2)
    s = D ** -0.25
    w = torch.einsum("bthd,bshd->bhts", q * s, k * s).softmax(-1)
    return torch.einsum("bhts,bshd->bthd", w, v).reshape(B, T, Cc)


ATT_BWD_CASES = [
    # B, T, C, heads, order
    (2, 64, 128, 2, 0),       # head_dim 64, one tile
    (1, 256, 128, 2, 1),      # new attention order
    (2, 100, 64, 2, 0),       # head_dim 32, ragged T
    (1, 200, 64, 4, 1),       # head_dim 16, ragged
    (1, 1024, 256, 4, 0),     # 16 tiles
]


@pytest.mark.parametrize("case", ATT_BWD_CASES)
def test_attention_bwd_kernel(be, case):
    B, T, Cc, heads, order = case
    qkv = rnd((B, T, 3 * Cc), 30, 1.5)
    dout = rnd((B, T, Cc), 31, 0.3)
    qd = qkv.double().requires_grad_(True)
    od = _attention_ref(qd, heads, order)
    od.backward(dout.double())
    out = od.detach().float().to(DEV)
    dqkv = torch.full((B, T, 3 * Cc), float("nan"), device=DEV)
    lse, delta = torch.empty(B * heads * T, device=DEV), torch.empty(B * heads * T, device=DEV)
    be.attention_bwd(qkv.to(DEV), out, dout.to(DEV), heads, order, dqkv, lse, delta)
    torch.cuda.synchronize()
    assert not torch.isnan(dqkv).any()
    assert rel_dev(dqkv, qd.grad) < 2e-5, rel_dev(dqkv, qd.grad)


@pytest.mark.parametrize("B,H,W,C,heads,order", [(2, 8, 8, 128, 2, 0), (1, 16, 16, 64, 2, 1), (2, 16, 16, 64, 4, 0)])
def test_attention_core_function(B, H, W, C, heads, order):
    """AttentionCoreFn (native forward kernels + flash backward) vs the fp64 torch graph."""
    from bbdm_b200.train import AttentionCoreFn
    qkv = (rnd((B, 3 * C, H, W), 32, 1.2).to(DEV).contiguous(memory_format=torch.channels_last)).requires_grad_(True)
    gy = rnd((B, C, H, W), 33, 0.3).to(DEV)
    y = AttentionCoreFn.apply(qkv, heads, order)
    y.backward(gy)
    qd = qkv.detach().double().cpu().requires_grad_(True)
    od = _attention_ref(qd.permute(0, 2, 3, 1).reshape(B, H * W, 3 * C), heads, order)
    od.backward(gy.double().cpu().permute(0, 2, 3, 1).reshape(B, H * W, C))
    assert rel_dev(y.permute(0, 2, 3, 1).reshape(B, H * W, C), od) < 3e-5
    assert rel_dev(qkv.grad, qd.grad) < 5e-5, rel_dev(qkv.grad, qd.grad)


def test_attention_block_training_matches_torch_graph():
    """AttentionBlock.forward in training: native GN+qkv, attention core, proj vs the stock-PyTorch path."""
    import bbdm_b200.unet as U
    blk = U.AttentionBlock(128, num_head_channels=64).to(DEV)
    with torch.no_grad():
        for p_ in blk.parameters():
            p_.copy_(rnd(tuple(p_.shape), 40 + p_.numel() % 7, 0.05).to(DEV))
        blk.norm.weight.add_(1.0)
    x = rnd((2, 128, 16, 16), 41).to(DEV)
    gy = rnd((2, 128, 16, 16), 42, 0.2).to(DEV)
    res = {}
    for native in (True, False):
        U.NATIVE_TRAIN_CONV = native
        blk.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        y = blk(xi)
        y.backward(gy)
        res[native] = (y.detach(), xi.grad, {n: p_.grad.clone() for n, p_ in blk.named_parameters()})
    U.NATIVE_TRAIN_CONV = True
    assert rel_dev(res[True][0], res[False][0]) < 3e-5
    assert rel_dev(res[True][1], res[False][1]) < 1e-4
    for n in res[False][2]:
        assert rel_dev(res[True][2][n], res[False][2][n]) < 1e-4, n


@pytest.mark.parametrize("Cout,Cin,k", [(64, 64, 3), (128, 192, 3), (256, 64, 1), (96, 40, 3), (33, 65, 1)])
def test_pack_weight_split_both_equals_single_layout_packers(be, Cout, Cin, k):
    w = rnd((Cout, Cin, k, k), 50, 0.05).to(DEV)
    mk = lambda *s: torch.full(s, 7.0, dtype=torch.bfloat16, device=DEV)
    f_hi, f_lo, d_hi, d_lo = mk(k * k, Cout, Cin), mk(k * k, Cout, Cin), mk(k * k, Cin, Cout), mk(k * k, Cin, Cout)
    r_hi, r_lo, s_hi, s_lo = mk(k * k, Cout, Cin), mk(k * k, Cout, Cin), mk(k * k, Cin, Cout), mk(k * k, Cin, Cout)
    be.pack_weight_split_both(w, f_hi, f_lo, d_hi, d_lo)
    be.pack_weight_split(w, r_hi, r_lo)
    be.pack_weight_split_dgrad(w, s_hi, s_lo)
    for a, b in ((f_hi, r_hi), (f_lo, r_lo), (d_hi, s_hi), (d_lo, s_lo)):
        assert torch.equal(a, b)
    g_hi, g_lo = mk(k * k, Cout, Cin), mk(k * k, Cout, Cin)
    be.pack_weight_split_both(w, g_hi, g_lo)                 # forward layout only
    assert torch.equal(g_hi, r_hi) and torch.equal(g_lo, r_lo)
