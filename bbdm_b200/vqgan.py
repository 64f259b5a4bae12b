"""Parameter tree of the VQGAN autoencoder with the reference's state_dict names and constructor surface
(model/VQGAN/vqgan.py:30-59, model.py:342-537, quantize.py:213-246), so reference ``.ckpt`` files load and
``VQGANEngine`` can walk either this tree or the reference's own ``VQModel``.

These modules are CONTAINERS: the computation is ``bbdm_b200.vqgan_engine.VQGANEngine`` on the sm_100a
kernels.  The latent model keeps using the reference's frozen ``VQModel`` when the BBDM checkout is on the
path (north star); this tree serves deployments and tests without it (e.g. the GPU test box).  There is no
CPU arithmetic here -- calling it on CPU tensors raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def _norm(c):
    return nn.GroupNorm(num_groups=32, num_channels=c, eps=1e-6, affine=True)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, 3, 1, 1)


class Downsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, 3, 2, 0)      # input zero-padded (0,1,0,1) first


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=0):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels, self.use_conv_shortcut = in_channels, out_channels, conv_shortcut
        self.norm1 = _norm(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        if temb_channels > 0:
            self.temb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = _norm(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
            else:
                self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)


class AttnBlock(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = _norm(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, 1)
        self.k = nn.Conv2d(in_channels, in_channels, 1)
        self.v = nn.Conv2d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)


def _mid(block_in):
    mid = nn.Module()
    mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in)
    mid.attn_1 = AttnBlock(block_in)
    mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in)
    return mid


class Encoder(nn.Module):
    def __init__(self, *, ch, out_ch=None, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, **ignore_kwargs):
        super().__init__()
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        curr_res, in_ch_mult = resolution, (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i], ch * ch_mult[i]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res //= 2
            self.down.append(down)
        self.mid = _mid(block_in)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, 1, 1)


class Decoder(nn.Module):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels=None, resolution, z_channels, give_pre_end=False, **ignore_kwargs):
        super().__init__()
        if give_pre_end:
            raise NotImplementedError("give_pre_end")
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.resolution = resolution
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = _mid(block_in)
        self.up = nn.ModuleList()
        for i in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res *= 2
            self.up.insert(0, up)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)


class VectorQuantizer(nn.Module):
    def __init__(self, n_e, e_dim):
        super().__init__()
        self.n_e, self.e_dim = n_e, e_dim
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)


class VQModel(nn.Module):
    """Same ctor keys as the reference (``VQGAN.params`` of Template-LBBDM-*.yaml); ``lossconfig`` is accepted
    and ignored (the autoencoder is frozen here, its training loss is out of scope)."""

    def __init__(self, ddconfig, lossconfig=None, n_embed=None, embed_dim=None, ckpt_path=None, ignore_keys=(),
                 image_key="image", colorize_nlabels=None, monitor=None, remap=None, sane_index_shape=False):
        super().__init__()
        if remap is not None:
            raise NotImplementedError("VectorQuantizer remap")
        dd = vars(ddconfig) if not isinstance(ddconfig, dict) else ddconfig
        self.encoder = Encoder(**dd)
        self.decoder = Decoder(**dd)
        self.quantize = VectorQuantizer(n_embed, embed_dim)
        self.quant_conv = nn.Conv2d(dd["z_channels"], embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, dd["z_channels"], 1)
        self._engine = None
        if ckpt_path is not None:
            sd = torch.load(ckpt_path, map_location="cpu")["state_dict"]
            sd = {k: v for k, v in sd.items() if not any(k.startswith(ik) for ik in ignore_keys)}
            self.load_state_dict(sd, strict=False)
            print(f"Restored from {ckpt_path}")

    def engine(self):
        if self._engine is None:
            from .vqgan_engine import VQGANEngine
            self._engine = VQGANEngine(self)
        return self._engine

    @torch.no_grad()
    def encode_latent(self, x, quant_conv=True):
        return self.engine().encode(x, quant_conv=quant_conv)

    @torch.no_grad()
    def decode_latent(self, z, quant_conv_first=False):
        return self.engine().decode(z, quant_conv_first=quant_conv_first)
