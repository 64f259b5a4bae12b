"""Host-side logic of the UNet executor, checked on CPU against the reference-generated
fixtures through an oracle-backed emulation of the kernel backend (tests/_emu_backend.py).
The CUDA kernels themselves are checked by the -m gpu suite."""
import os

import numpy as np
import pytest
import torch

from _emu_backend import EmuBackend
from _recipe import UNET_CONFIGS, fill_state_dict, rel_dev
from bbdm_b200.engine import UNetEngine
from bbdm_b200.unet import UNetModel

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def build(unet_name):
    net = UNetModel(**UNET_CONFIGS[unet_name]).eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(fill_state_dict(shapes, seed=1234))
    return net


@pytest.mark.parametrize("tag,tol,expect_umma", [
    ("tiny_pixel", 2e-5, True),      # 256-channel level runs through the tensor-core conv path
    ("tiny_latent", 2e-5, True),
    ("tiny_variant", 2e-5, True),    # no scale-shift norm, conv up/down, new attention order
    ("mid_pixel", 6e-5, True),       # aligned channels: split-bf16 operand planes (2^-17 rounding)
    ("tiny_st", 6e-5, True),         # SpatialTransformer blocks: LayerNorm / GEGLU / self- and cross-attention
])
def test_engine_wiring_matches_reference_fixture(tag, tol, expect_umma):
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, tag + ".npz")).items() if v.ndim}
    net = build(tag)
    be = EmuBackend()
    eng = UNetEngine(net, backend=be)
    ctx = None if net.condition_key == "nocond" else g["y"]
    out = eng.forward(g["x"], g["t"], ctx)
    assert out.shape == g["unet_out"].shape and not torch.isnan(out).any()
    assert rel_dev(out, g["unet_out"]) < tol
    assert ("conv_umma" in be.calls) == expect_umma
    if tag == "tiny_st":
        assert {"layernorm_split", "geglu_split", "attention_cross"} <= set(be.calls)
    # second call: no new allocations (stable addresses for CUDA-graph replay), same result
    pool = eng._pool(g["x"].device, tuple(g["x"].shape[i] for i in (0, 2, 3)))
    nbytes = pool.bytes
    out2 = eng.forward(g["x"], g["t"], ctx)
    assert pool.bytes == nbytes
    assert torch.equal(out, out2)


def test_winograd_path_wiring_matches_reference_fixture():
    """The Winograd F(4x4,3x3) route of the ResBlock convs (input transform -> 36 position GEMMs -> output transform,
    1x1 skip as a separate GEMM entering as the residual, GN partial sums from the output transform), forced on for
    the 64-channel 32x32 level of mid_pixel (B*8*8 = 128 tiles) through the emulation backend."""
    g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, "mid_pixel.npz")).items() if v.ndim}
    net = build("mid_pixel")
    be = EmuBackend()
    eng = UNetEngine(net, backend=be)
    assert eng.wino
    eng.wino_min_c, eng.wino_min_tiles = 64, 128
    out = eng.forward(g["x"], g["t"], g["y"])
    assert be.calls.count("wino_input") >= 4 and be.calls.count("wino_input") == be.calls.count("wino_output")
    assert not torch.isnan(out).any()
    assert rel_dev(out, g["unet_out"]) < 6e-5
    eng2 = UNetEngine(net, backend=EmuBackend())
    eng2.wino = False
    assert rel_dev(out, eng2.forward(g["x"], g["t"], g["y"])) < 8e-5     # both are within 6e-5 of the reference


def test_weight_cache_refresh_on_param_change():
    net = build("tiny_latent")
    be = EmuBackend()
    eng = UNetEngine(net, backend=be)
    eng.refresh_weights()
    n0 = be.calls.count("pack_weight_f32")
    eng.refresh_weights()
    assert be.calls.count("pack_weight_f32") == n0            # unchanged -> cached
    with torch.no_grad():
        net.out[2].weight.add_(1.0)                             # optimizer-style in-place update
    eng.refresh_weights()
    assert be.calls.count("pack_weight_f32") == 2 * n0
    p = net.out[2].weight
    gen = eng.generation
    addr = {k: v["f32"].data_ptr() for k, v in eng._w.items() if isinstance(v, dict) and "f32" in v}
    p.data = p.data.clone()                                     # EMA-style .data swap
    eng.refresh_weights()
    assert be.calls.count("pack_weight_f32") == 3 * n0
    # the derived caches are re-packed into the same buffers (no reallocation per EMA swap); captured graphs are
    # invalidated because biases / norm affines are read through the parameters' own (changed) addresses
    assert {k: v["f32"].data_ptr() for k, v in eng._w.items() if isinstance(v, dict) and "f32" in v} == addr
    assert eng.generation == gen + 1


def test_unet_rejects_cpu_inference():
    net = build("tiny_latent")
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 4, 16, 16), timesteps=torch.zeros(1, dtype=torch.long))


def test_buffer_pools_are_bounded_across_shape_changes():
    net = build("tiny_latent")
    eng = UNetEngine(net, backend=EmuBackend())
    t = torch.zeros(1, dtype=torch.long)
    for b in (1, 2, 3, 2, 1):
        eng.forward(torch.zeros(b, 4, 16, 16), t.expand(b).contiguous(), None)
        assert len(eng._pools) <= 2
    assert eng.pool_bytes() > 0
