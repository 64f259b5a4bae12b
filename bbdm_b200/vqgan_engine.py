"""VQGAN encode / decode executor over the C-ABI kernels.

The latent models call the frozen autoencoder at both ends of the bridge loop and twice per training
sample (LatentBrownianBridgeModel.py:57-100 of the reference).  This executor walks the reference's own
``VQModel`` module tree (model/VQGAN/vqgan.py:30-88, model.py:342-537 -- the module stays the parameter
container, exactly like the UNet) and issues the same kernels as the UNet executor:

  ResnetBlock (model.py:76-138)   stats -> prep (GN eps 1e-6 + swish + split) -> tcgen05 conv, twice; the 1x1
                                  nin_shortcut rides as extra K-blocks of conv2, the identity skip as its residual
  AttnBlock   (model.py:140-192)  single head of width C: C <= 64 -> the flash kernels; C >= 128 -> per image two
                                  tensor-core GEMMs (S = Q K^T, O = P V) around bbdm_softmax_rows_split
  Downsample  (model.py:55-73)    zero-pad (0,1,0,1) + stride-2 conv = space-to-depth split + 2x2-tap tcgen05 conv
                                  (bbdm_conv_direct_pad for unaligned channel counts)
  Upsample    (model.py:38-53)    nearest-2x + conv3x3 = the fused 4-phase tcgen05 conv (no upsampled tensor)
  VectorQuantizer2 (quantize.py:271-312)  bbdm_vq_nearest

Inference only (the autoencoder is frozen and always called under no_grad).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import cabi
from .engine import GN_GROUPS, KernelExecutor
from .weights import upsample_phase_weights


class VQGANEngine(KernelExecutor):
    gn_eps = 1e-6                      # model/VQGAN/model.py:34-35

    def __init__(self, vqmodel: nn.Module, backend=None, precision: str = "split3"):
        super().__init__(backend, precision)
        self.vq = vqmodel
        self._w = {}
        self._wkey = None

    # ------------------------------------------------------------------------------ weights
    def _params_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.vq.parameters())

    def refresh_weights(self, force=False):
        key = self._params_key()
        if not force and key == self._wkey:
            return
        be = self.be
        dev = next(self.vq.parameters()).device
        w = {}

        def pack(name, wt, bias):
            wt = wt.detach().contiguous()
            cout, cin, k = wt.shape[0], wt.shape[1], wt.shape[2]
            ent = {"cout": cout, "cin": cin, "k": k, "bias": None if bias is None else bias.detach().contiguous()}
            if cin % 64 == 0 and cout % 64 == 0 and k in (1, 3):
                ent["hi"] = be.empty((k * k, cout, cin), torch.bfloat16, dev)
                ent["lo"] = be.empty((k * k, cout, cin), torch.bfloat16, dev)
                be.pack_weight_split(wt, ent["hi"], ent["lo"])
            elif cin % 64 == 0 and cout < 64 and k == 3:
                # image head (Cout = 3): zero-padded to one 64-wide N tile, epilogue stores NCHW
                ent["hi_pad"] = torch.zeros((k * k, 64, cin), dtype=torch.bfloat16, device=dev)
                ent["lo_pad"] = torch.zeros((k * k, 64, cin), dtype=torch.bfloat16, device=dev)
                ent["bias_pad"] = torch.zeros((64,), dtype=torch.float32, device=dev)
                if bias is not None:
                    ent["bias_pad"][:cout].copy_(bias.detach())
                be.pack_weight_split(wt, ent["hi_pad"], ent["lo_pad"])
            ent["f32"] = be.empty((k * k, cin, cout), torch.float32, dev)
            be.pack_weight_f32(wt, ent["f32"])
            if self.wino and "hi" in ent and k == 3 and min(cin, cout) >= self.wino_min_c \
                    and (name.endswith(".conv1") or name.endswith(".conv2")):
                # ResnetBlock 3x3 convs: Winograd-domain planes (csrc/winograd.cu), like the UNet's ResBlocks
                ent["u_hi"] = be.empty((36, cout, cin), torch.float16, dev)
                ent["u_lo"] = be.empty((36, cout, cin), torch.float16, dev)
                be.wino_pack_weight(wt, ent["u_hi"], ent["u_lo"])
            w[name] = ent

        for name, m in self.vq.named_modules():
            if isinstance(m, nn.Conv2d) and not name.startswith("loss"):
                pack(name, m.weight, m.bias)
            if type(m).__name__ == "AttnBlock" and m.in_channels <= 64:
                # q | k | v as one 1x1 conv ("new order" layout of the flash kernels, one head)
                pack(name + ".qkv", torch.cat([m.q.weight, m.k.weight, m.v.weight], 0),
                     torch.cat([m.q.bias, m.k.bias, m.v.bias], 0))
        for name, m in self.vq.named_modules():          # second pass: the convs above are packed now
            if type(m).__name__ == "Upsample" and m.with_conv and "hi" in w.get(name + ".conv", {}):
                ent = w[name + ".conv"]
                wp = upsample_phase_weights(m.conv.weight.detach())
                ent["up_hi"] = be.empty((16, ent["cout"], ent["cin"]), torch.bfloat16, dev)
                ent["up_lo"] = be.empty((16, ent["cout"], ent["cin"]), torch.bfloat16, dev)
                be.pack_weight_split_taps(wp, ent["up_hi"], ent["up_lo"])
            if type(m).__name__ == "Downsample" and m.with_conv:
                # stride-2 3x3 conv on the zero-padded input == 2x2-tap conv over the space-to-depth tensor:
                # W2[tap=(ty,tx)][co][(a*2+b)*C + ci] = w[co][ci][2ty+a][2tx+b]  (zero where 2ty+a or 2tx+b = 3)
                cw = m.conv.weight.detach()
                co, ci = cw.shape[0], cw.shape[1]
                if co % 64 == 0 and (4 * ci) % 64 == 0:
                    w2 = torch.zeros((co, 4, ci, 4), dtype=torch.float32, device=dev)     # [co][a*2+b][ci][tap]
                    for ty in range(2):
                        for tx in range(2):
                            for a in range(2):
                                for b in range(2):
                                    if 2 * ty + a < 3 and 2 * tx + b < 3:
                                        w2[:, a * 2 + b, :, ty * 2 + tx] = cw[:, :, 2 * ty + a, 2 * tx + b]
                    ent = w[name + ".conv"]
                    ent["ds_hi"] = be.empty((4, co, 4 * ci), torch.bfloat16, dev)
                    ent["ds_lo"] = be.empty((4, co, 4 * ci), torch.bfloat16, dev)
                    be.pack_weight_split_taps(w2.reshape(co, 4 * ci, 4).contiguous(), ent["ds_hi"], ent["ds_lo"])
        self._w, self._wkey = w, key

    # ------------------------------------------------------------------------------ pieces
    def _split(self, pool, x):
        hi, lo = pool.get(x.shape, torch.bfloat16), pool.get(x.shape, torch.bfloat16)
        self.be.prep(x, None, raw_hi=hi, raw_lo=lo)
        return hi, lo

    def _conv_plain(self, pool, ent, x, **kw):
        """conv on an fp32 NHWC tensor with no normalisation in front (conv_in, quant convs, shortcuts)."""
        B, H, W, _ = x.shape
        if "hi" in ent and W >= 4:
            hi, lo = self._split(pool, x)
            out, _, _ = self._conv(pool, ent, a_hi=hi, a_lo=lo, shape=(B, H, W), **kw)
            pool.put(hi, lo)
            return out
        out, _, _ = self._conv(pool, ent, a_f32=x, shape=(B, H, W), **kw)
        return out

    def _gn_act(self, pool, x, norm, umma, silu=True, want_raw_split=False):
        """GroupNorm(eps 1e-6) (+ swish) of x as a conv operand: (a_f32, a_hi, a_lo, raw_hi, raw_lo)."""
        mean, rstd = self._stats(pool, x, None)
        a_f32 = a_hi = a_lo = r_hi = r_lo = None
        if umma:
            a_hi, a_lo = pool.get(x.shape, torch.bfloat16), pool.get(x.shape, torch.bfloat16)
        else:
            a_f32 = pool.get(x.shape)
        if want_raw_split:
            r_hi, r_lo = pool.get(x.shape, torch.bfloat16), pool.get(x.shape, torch.bfloat16)
        self.be.prep(x, None, groups=GN_GROUPS, mean=mean, rstd=rstd, gamma=norm.weight.detach(),
                     beta=norm.bias.detach(), silu=silu, resample=cabi.RESAMPLE_NONE, act_f32=a_f32, act_hi=a_hi,
                     act_lo=a_lo, raw_hi=r_hi, raw_lo=r_lo)
        pool.put(mean, rstd)
        return a_f32, a_hi, a_lo, r_hi, r_lo

    def _resnet(self, pool, name, m, x):
        w = self._w
        B, H, W, cin = x.shape
        e1, e2 = w[name + ".conv1"], w[name + ".conv2"]
        es = w.get(name + ".nin_shortcut", w.get(name + ".conv_shortcut"))
        assert (es is None) == (m.in_channels == m.out_channels)
        umma1, umma2 = "hi" in e1 and W >= 4, "hi" in e2 and W >= 4
        fuse_skip = es is not None and es["k"] == 1 and umma2 and "hi" in es
        skip_umma = es is not None and not fuse_skip and "hi" in es and W >= 4
        wino1, wino2 = umma1 and self._wino_ok(e1, B, H, W), umma2 and self._wino_ok(e2, B, H, W)
        r_hi = r_lo = None
        if wino1:
            # Winograd conv1: the raw split planes a 1x1 shortcut needs come out of the same input pass
            if fuse_skip or skip_umma:
                r_hi, r_lo = pool.get(x.shape, torch.bfloat16), pool.get(x.shape, torch.bfloat16)
            mean, rstd = self._stats(pool, x, None)
            h1 = self._wino_conv(pool, e1, x, None, mean=mean, rstd=rstd, gamma=m.norm1.weight.detach(),
                                 beta=m.norm1.bias.detach(), raw=None if r_hi is None else (r_hi, r_lo))
            pool.put(mean, rstd)
        else:
            a_f32, a_hi, a_lo, r_hi, r_lo = self._gn_act(pool, x, m.norm1, umma1, want_raw_split=fuse_skip or skip_umma)
            h1, _, _ = self._conv(pool, e1, a_f32=a_f32, a_hi=a_hi, a_lo=a_lo, shape=(B, H, W), stats=True)
            pool.put(a_f32, a_hi, a_lo)
        if wino2:
            # Winograd conv2: the shortcut (1x1 GEMM or identity) enters as the output transform's residual
            sk = None
            if es is not None:
                if "hi" in es and W >= 4 and r_hi is not None:
                    sk, _, _ = self._conv(pool, es, a_hi=r_hi, a_lo=r_lo, shape=(B, H, W))
                else:
                    sk, _, _ = self._conv(pool, es, a_f32=x, shape=(B, H, W))
            mean, rstd = self._stats(pool, h1, None)
            out = self._wino_conv(pool, e2, h1, None, mean=mean, rstd=rstd, gamma=m.norm2.weight.detach(),
                                  beta=m.norm2.bias.detach(), residual=x if sk is None else sk, res_mode=cabi.RES_SAME)
            pool.put(mean, rstd, h1, r_hi, r_lo, sk)
            return out
        b_f32, b_hi, b_lo, _, _ = self._gn_act(pool, h1, m.norm2, umma2)
        pool.put(h1)
        kw = dict(a_f32=b_f32, a_hi=b_hi, a_lo=b_lo, shape=(B, H, W), stats=True)
        sk = None
        if fuse_skip:
            out, _, _ = self._conv(pool, e2, second=(es, r_hi, r_lo), **kw)
        elif es is not None:
            if skip_umma:
                sk, _, _ = self._conv(pool, es, a_hi=r_hi, a_lo=r_lo, shape=(B, H, W))
            else:
                sk, _, _ = self._conv(pool, es, a_f32=x, shape=(B, H, W))
            out, _, _ = self._conv(pool, e2, residual=sk, res_mode=cabi.RES_SAME, **kw)
        else:
            out, _, _ = self._conv(pool, e2, residual=x, res_mode=cabi.RES_SAME, **kw)
        pool.put(b_f32, b_hi, b_lo, r_hi, r_lo, sk)
        return out

    def _attn(self, pool, name, m, x):
        be, w = self.be, self._w
        B, H, W, Cc = x.shape
        T = H * W
        ep = w[name + ".proj_out"]
        umma = self._umma_ok(Cc, Cc, W)
        a_f32, a_hi, a_lo, _, _ = self._gn_act(pool, x, m.norm, umma, silu=False)
        o_f32 = o_hi = o_lo = None
        if Cc in (16, 32, 64):
            # one head of width C: the flash kernel's scale D^-1/4 on q and on k is the reference's C^-1/2
            qkv, _, _ = self._conv(pool, w[name + ".qkv"], a_f32=a_f32, a_hi=a_hi, a_lo=a_lo, shape=(B, H, W))
            if umma:
                o_hi, o_lo = pool.get(x.shape, torch.bfloat16), pool.get(x.shape, torch.bfloat16)
            else:
                o_f32 = pool.get(x.shape)
            be.attention(qkv.view(B, T, 3 * Cc), 1, 1, None if o_f32 is None else o_f32.view(B, T, Cc),
                         None if o_hi is None else o_hi.view(B, T, Cc), None if o_lo is None else o_lo.view(B, T, Cc))
            pool.put(qkv)
        else:
            if not (umma and T % 64 == 0):
                raise NotImplementedError(f"VQGAN AttnBlock with C={Cc}, T={T}: needs C % 64 == 0 and T % 64 == 0")
            _, q_hi, q_lo = self._conv(pool, w[name + ".q"], a_hi=a_hi, a_lo=a_lo, shape=(B, H, W), out_split=True,
                                       want_f32=False)
            _, k_hi, k_lo = self._conv(pool, w[name + ".k"], a_hi=a_hi, a_lo=a_lo, shape=(B, H, W), out_split=True,
                                       want_f32=False)
            v, _, _ = self._conv(pool, w[name + ".v"], a_hi=a_hi, a_lo=a_lo, shape=(B, H, W))
            vt_hi, vt_lo = pool.get((B, Cc, T), torch.bfloat16), pool.get((B, Cc, T), torch.bfloat16)
            o_hi, o_lo = pool.get(x.shape, torch.bfloat16), pool.get(x.shape, torch.bfloat16)
            s = pool.get((1, H, W, T))
            p_hi, p_lo = pool.get((1, H, W, T), torch.bfloat16), pool.get((1, H, W, T), torch.bfloat16)
            scale = float(int(Cc) ** (-0.5))
            for b in range(B):
                # V^T planes [C][T] = the K-major B operand of O = P V (same kernel as the wgrad operand split)
                be.split_grad(v[b].view(T, Cc), None, None, vt_hi[b], vt_lo[b])
                # S[t, s] = sum_c q[t, c] k[s, c]: the K planes of this image ARE a [Cout=T][Cin=C] weight
                be.conv_umma(B=1, H=H, W=W, Cin=Cc, Cout=T, taps=1, a_hi=q_hi[b:b + 1], a_lo=q_lo[b:b + 1],
                             w_hi=k_hi[b].view(1, T, Cc), w_lo=k_lo[b].view(1, T, Cc), out=s, passes=self.passes)
                be.softmax_rows_split(s.view(T, T), scale, p_hi.view(T, T), p_lo.view(T, T))
                be.conv_umma(B=1, H=H, W=W, Cin=T, Cout=Cc, taps=1, a_hi=p_hi, a_lo=p_lo,
                             w_hi=vt_hi[b].view(1, Cc, T), w_lo=vt_lo[b].view(1, Cc, T), out=None,
                             out_hi=o_hi[b:b + 1], out_lo=o_lo[b:b + 1], passes=self.passes)
            pool.put(q_hi, q_lo, k_hi, k_lo, v, vt_hi, vt_lo, s, p_hi, p_lo)
        pool.put(a_f32, a_hi, a_lo)
        out, _, _ = self._conv(pool, ep, a_f32=o_f32, a_hi=o_hi, a_lo=o_lo, shape=(B, H, W), residual=x,
                               res_mode=cabi.RES_SAME, stats=True)
        pool.put(o_f32, o_hi, o_lo)
        return out

    def _downsample(self, pool, name, m, x):
        B, H, W, Cc = x.shape
        if m.with_conv:
            ent = self._w[name + ".conv"]
            if "ds_hi" in ent and H % 2 == 0 and W % 2 == 0 and W // 2 >= 4:
                hi = pool.get((B, H // 2, W // 2, 4 * Cc), torch.bfloat16)
                lo = pool.get((B, H // 2, W // 2, 4 * Cc), torch.bfloat16)
                self.be.s2d_split(x, hi, lo)
                out = pool.get((B, H // 2, W // 2, Cc))
                rows = self._geom(H // 2, W // 2)
                part = pool.get((B * rows, Cc, 2)) if rows else None
                self.be.conv_umma(B=B, H=H // 2, W=W // 2, Cin=4 * Cc, Cout=Cc, taps=4, a_hi=hi, a_lo=lo,
                                  w_hi=ent["ds_hi"], w_lo=ent["ds_lo"], bias=ent["bias"], out=out, passes=self.passes,
                                  stats_partial=part)
                if part is not None:
                    out._gn = (part, rows)
                pool.put(hi, lo)
                return out
            out = pool.get((B, (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1, Cc))
            self.be.conv_direct_pad(x, ent["f32"], ent["bias"], None, out, Cc, 3, 2, 0, 1)
            return out
        out = pool.get((B, H // 2, W // 2, Cc))
        self.be.prep(x, None, resample=cabi.RESAMPLE_DOWN2, raw_f32=out)
        return out

    def _upsample(self, pool, name, m, x):
        be = self.be
        B, H, W, Cc = x.shape
        if not m.with_conv:
            up = pool.get((B, 2 * H, 2 * W, Cc))
            be.prep(x, None, resample=cabi.RESAMPLE_UP2, raw_f32=up)
            return up
        ent = self._w[name + ".conv"]
        if "up_hi" in ent and W >= 4:
            hi, lo = self._split(pool, x)
            out = pool.get((B, 2 * H, 2 * W, Cc))
            rows = 4 * self._geom(H, W)
            part = pool.get((B * rows, Cc, 2)) if rows else None
            be.conv_umma(B=B, H=H, W=W, Cin=Cc, Cout=Cc, taps=4, a_hi=hi, a_lo=lo, w_hi=ent["up_hi"], w_lo=ent["up_lo"],
                         bias=ent["bias"], out=out, passes=self.passes, upsample2x=True, stats_partial=part)
            if part is not None:
                out._gn = (part, rows)
            pool.put(hi, lo)
            return out
        up = pool.get((B, 2 * H, 2 * W, Cc))
        be.prep(x, None, resample=cabi.RESAMPLE_UP2, raw_f32=up)
        out, _, _ = self._conv(pool, ent, a_f32=up, shape=(B, 2 * H, 2 * W))
        pool.put(up)
        return out

    def _step(self, pool, h, fn, *args):
        new = fn(pool, *args, h)
        pool.put(h)
        return new

    def _mid(self, pool, prefix, mid, h):
        h = self._step(pool, h, self._resnet, prefix + ".block_1", mid.block_1)
        h = self._step(pool, h, self._attn, prefix + ".attn_1", mid.attn_1)
        return self._step(pool, h, self._resnet, prefix + ".block_2", mid.block_2)

    def _to_nhwc(self, pool, x):
        B, Cx, H, W = x.shape
        xin = pool.get((B, H, W, Cx))
        self.be.nchw_to_nhwc_cat(x.contiguous().float(), None, xin)
        return xin

    def _head(self, pool, h, norm, ent, out_channels):
        """norm_out -> swish -> conv_out; returns (nhwc or None, nchw or None)."""
        B, H, W, _ = h.shape
        if "hi_pad" in ent and W >= 4:
            _, a_hi, a_lo, _, _ = self._gn_act(pool, h, norm, True)
            out = torch.empty((B, out_channels, H, W), dtype=torch.float32, device=h.device)
            self.be.conv_umma(B=B, H=H, W=W, Cin=ent["cin"], Cout=64, taps=9, a_hi=a_hi, a_lo=a_lo, w_hi=ent["hi_pad"],
                              w_lo=ent["lo_pad"], bias=ent["bias_pad"], out=out, passes=self.passes,
                              out_nchw_channels=out_channels)
            pool.put(a_hi, a_lo)
            return None, out
        umma = "hi" in ent and W >= 4
        a_f32, a_hi, a_lo, _, _ = self._gn_act(pool, h, norm, umma)
        y, _, _ = self._conv(pool, ent, a_f32=a_f32, a_hi=a_hi, a_lo=a_lo, shape=(B, H, W))
        pool.put(a_f32, a_hi, a_lo)
        return y, None

    # ------------------------------------------------------------------------------ public
    # activation budget of one pass: 32 images of 256x256 (the cfg3 batch; ~25 GB of pooled NHWC tensors and operand
    # planes).  Larger batches / resolutions run in chunks of this many image pixels -- the ends are 2-13 % of a
    # sampled batch, so nothing is lost, and BASELINE configs[3] (64 x 512^2) would otherwise need > 180 GB.
    max_pixels_per_pass = 32 * 256 * 256

    def _chunks(self, n_images, image_pixels):
        per = max(1, self.max_pixels_per_pass // max(1, image_pixels))
        return [(i, min(n_images, i + per)) for i in range(0, n_images, per)]

    @torch.no_grad()
    def encode(self, x, quant_conv=True):
        """vqgan.encoder(x) [-> vqgan.quant_conv]   (LatentBrownianBridgeModel.py:73-82): NCHW in, NCHW out."""
        ch = self._chunks(x.shape[0], x.shape[2] * x.shape[3])
        if len(ch) == 1:
            return self._encode_pass(x, quant_conv)
        return torch.cat([self._encode_pass(x[a:b], quant_conv) for a, b in ch], 0)

    def _encode_pass(self, x, quant_conv=True):
        self.refresh_weights()
        enc, w = self.vq.encoder, self._w
        pool = self._pool(x.device, ("enc",) + tuple(x.shape))
        h = self._to_nhwc(pool, x)
        h = self._step(pool, h, lambda p, t: self._conv_plain(p, w["encoder.conv_in"], t))
        for i in range(enc.num_resolutions):
            lvl = enc.down[i]
            for j in range(enc.num_res_blocks):
                h = self._step(pool, h, self._resnet, f"encoder.down.{i}.block.{j}", lvl.block[j])
                if len(lvl.attn) > 0:
                    h = self._step(pool, h, self._attn, f"encoder.down.{i}.attn.{j}", lvl.attn[j])
            if i != enc.num_resolutions - 1:
                h = self._step(pool, h, self._downsample, f"encoder.down.{i}.downsample", lvl.downsample)
        h = self._mid(pool, "encoder.mid", enc.mid, h)
        y, _ = self._head_nhwc(pool, h, enc.norm_out, w["encoder.conv_out"])
        pool.put(h)
        if quant_conv:
            y = self._step(pool, y, lambda p, t: self._conv_plain(p, w["quant_conv"], t))
        B, H, W, Cz = y.shape
        out = torch.empty((B, Cz, H, W), dtype=torch.float32, device=x.device)
        self.be.nhwc_to_nchw(y, out)
        pool.put(y)
        return out

    def _head_nhwc(self, pool, h, norm, ent):
        B, H, W, _ = h.shape
        umma = "hi" in ent and W >= 4
        a_f32, a_hi, a_lo, _, _ = self._gn_act(pool, h, norm, umma)
        y, _, _ = self._conv(pool, ent, a_f32=a_f32, a_hi=a_hi, a_lo=a_lo, shape=(B, H, W))
        pool.put(a_f32, a_hi, a_lo)
        return y, None

    @torch.no_grad()
    def quantize(self, z_nhwc, pool):
        cb = self.vq.quantize.embedding.weight.detach()
        zq = pool.get(z_nhwc.shape)
        idx = pool.get(z_nhwc.shape[:3], torch.int64)
        self.be.vq_nearest(z_nhwc, cb, zq, idx)
        return zq, idx

    @torch.no_grad()
    def decode(self, z, quant_conv_first=False, return_indices=False):
        """[quant_conv ->] quantize -> post_quant_conv -> decoder   (LatentBrownianBridgeModel.py:84-100)."""
        f = 2 ** (self.vq.decoder.num_resolutions - 1)
        ch = self._chunks(z.shape[0], z.shape[2] * z.shape[3] * f * f)
        if len(ch) == 1:
            return self._decode_pass(z, quant_conv_first, return_indices)
        outs = [self._decode_pass(z[a:b], quant_conv_first, return_indices) for a, b in ch]
        if return_indices:
            return torch.cat([o[0] for o in outs], 0), torch.cat([o[1] for o in outs], 0)
        return torch.cat(outs, 0)

    def _decode_pass(self, z, quant_conv_first=False, return_indices=False):
        self.refresh_weights()
        dec, w = self.vq.decoder, self._w
        pool = self._pool(z.device, ("dec",) + tuple(z.shape))
        h = self._to_nhwc(pool, z)
        if quant_conv_first:
            h = self._step(pool, h, lambda p, t: self._conv_plain(p, w["quant_conv"], t))
        zq, idx = self.quantize(h, pool)
        pool.put(h)
        indices = idx.clone() if return_indices else None
        pool.put(idx)
        h = self._step(pool, zq, lambda p, t: self._conv_plain(p, w["post_quant_conv"], t))
        h = self._step(pool, h, lambda p, t: self._conv_plain(p, w["decoder.conv_in"], t))
        h = self._mid(pool, "decoder.mid", dec.mid, h)
        for i in reversed(range(dec.num_resolutions)):
            lvl = dec.up[i]
            for j in range(dec.num_res_blocks + 1):
                h = self._step(pool, h, self._resnet, f"decoder.up.{i}.block.{j}", lvl.block[j])
                if len(lvl.attn) > 0:
                    h = self._step(pool, h, self._attn, f"decoder.up.{i}.attn.{j}", lvl.attn[j])
            if i != 0:
                h = self._step(pool, h, self._upsample, f"decoder.up.{i}.upsample", lvl.upsample)
        ent = w["decoder.conv_out"]
        y, out = self._head(pool, h, dec.norm_out, ent, ent["cout"])
        pool.put(h)
        if out is None:
            B, H, W, Co = y.shape
            out = torch.empty((B, Co, H, W), dtype=torch.float32, device=z.device)
            self.be.nhwc_to_nchw(y, out)
            pool.put(y)
        return (out, indices) if return_indices else out
