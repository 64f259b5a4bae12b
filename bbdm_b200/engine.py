"""UNet executor over the C-ABI kernels.

Walks the reference-shaped module tree (bbdm_b200.unet) and issues one C-ABI call per fused
step.  Data layout inside the UNet: NHWC fp32 activations in HBM; every tensor-core conv reads
its A operand as a split-bf16 plane pair produced by the one-pass ``prep`` kernel
(GroupNorm-affine + FiLM + SiLU + up/down-sampling + concat, fused) and writes fp32 NHWC with
bias / 1x1-skip / residual fused in its epilogue.  See DESIGN.md section "Kernels".

The executor is written against a *backend* object (``cabi.CudaBackend`` -- the only product
backend).  tests/ inject an oracle-backed emulation of the same method set to verify this host
logic (block wiring, concat order, FiLM offsets, skip modes) on CPU; the product never does.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import cabi
from .weights import upsample_phase_weights
from .transformer import SpatialTransformer
from .unet import (AttentionBlock, Downsample, ResBlock, TimestepEmbedSequential, UNetModel,
                   Upsample, timestep_embedding)

GN_GROUPS = 32
GN_EPS = 1e-5


def resample_to_res(resample):
    return {cabi.RESAMPLE_NONE: cabi.RES_SAME, cabi.RESAMPLE_UP2: cabi.RES_UP2,
            cabi.RESAMPLE_DOWN2: cabi.RES_DOWN2}[resample]


class _Pool:
    """Shape-keyed free lists.  The forward pass is the same acquire/release sequence every
    call, so after the first call no allocation happens and every intermediate keeps a stable
    address (required for CUDA-graph replay, avoids allocator traffic)."""

    _serial = 0

    def __init__(self, backend, device):
        self.be, self.device, self.free = backend, device, {}
        self.bytes = 0
        _Pool._serial += 1
        self.serial = _Pool._serial          # identity of this set of buffers (captured graphs key on it)

    def get(self, shape, dtype=torch.float32):
        key = (tuple(shape), dtype)
        lst = self.free.get(key)
        if lst:
            return lst.pop()
        t = self.be.empty(tuple(shape), dtype, self.device)
        self.bytes += t.numel() * t.element_size()
        return t

    def put(self, *ts):
        for t in ts:
            if t is not None:
                gn = t.__dict__.pop("_gn", None)      # fused GroupNorm partial sums travel with the tensor
                if gn is not None:
                    self.put(gn[0])
                self.free.setdefault((tuple(t.shape), t.dtype), []).append(t)


class KernelExecutor:
    """What every executor over the C-ABI kernels shares: the backend, the precision mode, the buffer
    pools, GroupNorm statistics (fused partials or a stats pass) and the convolution dispatch."""

    gn_eps = GN_EPS

    def __init__(self, backend=None, precision: str = "split3"):
        self.be = backend if backend is not None else cabi.CudaBackend()
        assert precision in ("split3", "bf16")
        self.passes = 3 if precision == "split3" else 1
        self.precision = precision
        self._pools = {}
        self._gn_ws = None
        self._geom_cache = {}
        # Winograd F(4x4,3x3) for the stride-1 3x3 convs with at least this many input and output channels
        # (parity mode only; below that the transform traffic outweighs the 4x MAC saving).  BBDM_WINOGRAD=0 disables.
        self.wino = precision == "split3" and os.environ.get("BBDM_WINOGRAD", "1") != "0"
        self.wino_min_c = int(os.environ.get("BBDM_WINO_MIN_C", "256"))
        # ... and at least this many 4x4 tiles per launch: below it the 36 position GEMMs have too few M tiles each
        # (measured: cfg1, 256 tiles, graph replay 3.9 -> 4.4 ms with Winograd; cfg3, 2048 tiles, 20.2 -> 17.1 ms)
        self.wino_min_tiles = int(os.environ.get("BBDM_WINO_MIN_TILES", "512"))
        self._wino_geom = {}

    def _umma_ok(self, cin, cout, w):
        return cin % 64 == 0 and cout % 64 == 0 and w >= 4

    # ------------------------------------------------------------------------------ helpers
    def _pool(self, device, shape_key=None):
        """Intermediate-buffer pool for one (device, input shape).  At most two shapes stay resident (the
        steady batch and e.g. a smaller last batch of sample_to_eval); older pools are dropped so a
        long-lived model does not accumulate HBM across shape changes."""
        key = (device, shape_key)
        p = self._pools.pop(key, None)
        if p is None:
            p = _Pool(self.be, device)
            while len(self._pools) >= 2:
                self._pools.pop(next(iter(self._pools)))       # evict least recently used
        self._pools[key] = p                                   # (re)insert as most recent
        return p

    def pool_serial(self, device, shape_key):
        return self._pool(device, shape_key).serial

    def pool_bytes(self):
        return sum(p.bytes for p in self._pools.values())

    def _stats(self, pool, src1, src2, eps=None):
        eps = self.gn_eps if eps is None else eps
        B = src1.shape[0]
        mean, rstd = pool.get((B, GN_GROUPS)), pool.get((B, GN_GROUPS))
        g1 = getattr(src1, "_gn", None)
        g2 = None if src2 is None else getattr(src2, "_gn", None)
        if g1 is not None and (src2 is None or g2 is not None):
            # both tensors came out of the tensor-core conv: its epilogue already reduced them
            self.be.gn_finalize_partials(g1[0], g1[1], None if g2 is None else g2[0], 0 if g2 is None else g2[1],
                                         B, src1.shape[1] * src1.shape[2], GN_GROUPS, eps, mean, rstd)
            return mean, rstd
        if self._gn_ws is None or self._gn_ws.numel() < B * GN_GROUPS * cabi.GN_MAX_SLICES * 2 \
                or self._gn_ws.device != src1.device:
            self._gn_ws = self.be.empty((B * GN_GROUPS * cabi.GN_MAX_SLICES * 2,), torch.float64, src1.device)
        self.be.gn_stats(src1, src2, GN_GROUPS, eps, mean, rstd, self._gn_ws)
        return mean, rstd

    def _conv(self, pool, ent, *, a_f32=None, a_hi=None, a_lo=None, shape, bias=None, residual=None,
              res_mode=cabi.RES_NONE, second=None, out_split=False, want_f32=True, stride=1, out=None,
              stats=False):
        """One convolution.  shape = (B,H,W) of the INPUT; returns (out_f32, out_hi, out_lo)."""
        B, H, W = shape
        cout, cin, k = ent["cout"], ent["cin"], ent["k"]
        bias = ent["bias"] if bias is None else bias
        if a_hi is not None:
            if out is None:
                out = pool.get((B, H, W, cout)) if want_f32 else None
            oh = ol = None
            if out_split:
                oh, ol = pool.get((B, H, W, cout), torch.bfloat16), pool.get((B, H, W, cout), torch.bfloat16)
            kw = {}
            if second is not None:
                e2, r_hi, r_lo = second
                kw = dict(Cin2=e2["cin"], a2_hi=r_hi, a2_lo=r_lo, w2_hi=e2["hi"], w2_lo=e2["lo"], bias2=e2["bias"])
            part = None
            if stats and out is not None:
                rows = self._geom(H, W)
                if rows:
                    part = pool.get((B * rows, cout, 2))
            self.be.conv_umma(B=B, H=H, W=W, Cin=cin, Cout=cout, taps=k * k, a_hi=a_hi, a_lo=a_lo,
                              w_hi=ent["hi"], w_lo=ent["lo"], bias=bias, residual=residual, res_mode=res_mode,
                              out=out, out_hi=oh, out_lo=ol, passes=self.passes, stats_partial=part, **kw)
            if part is not None:
                out._gn = (part, rows)
            return out, oh, ol
        assert second is None and res_mode in (cabi.RES_NONE, cabi.RES_SAME) and not out_split
        Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
        if out is None:
            out = pool.get((B, Ho, Wo, cout))
        self.be.conv_direct(a_f32, ent["f32"], bias, residual, out, cout, k, stride)
        return out, None, None

    def _geom(self, H, W):
        key = (H, W)
        r = self._geom_cache.get(key)
        if r is None:
            r = self._geom_cache[key] = self.be.conv_geometry(H, W)[3]
        return r


    # ---- Winograd F(4x4,3x3) path (csrc/winograd.cu) ----------------------------------------------------------
    def _wino_geometry(self, B, H, W):
        key = (B, H, W)
        g = self._wino_geom.get(key)
        if g is None:
            g = self._wino_geom[key] = self.be.wino_geometry(B, H, W)
        return g

    def _wino_ok(self, ent, B, H, W):
        if not (self.wino and "u_hi" in ent):
            return False
        # >= 128 tiles per image: always (the choice must not depend on the batch size there -- batch-size independent,
        # bit-identical results at the pixel resolutions); smaller maps: only when the whole batch has enough tiles
        th, tw, tiles, ok = self._wino_geometry(B, H, W)
        return bool(ok and (th * tw >= 128 or tiles >= self.wino_min_tiles))

    def _wino_conv(self, pool, ent, src1, src2, *, mean, rstd, gamma, beta, film=None, silu=True, raw=None,
                   residual=None, res_mode=cabi.RES_NONE, stats=True):
        """GroupNorm-affine(+FiLM)+SiLU -> 3x3 conv (+bias, +residual) of cat(src1, src2) on the Winograd path:
        input transform -> 36 position GEMMs in one tcgen05 launch -> output transform (+ GN partial sums).
        raw = (r_hi, r_lo): also emit the raw input's split-bf16 planes (operand of a 1x1 skip conv)."""
        be = self.be
        B, H, W, _ = src1.shape
        cin, cout = ent["cin"], ent["cout"]
        th, tw, mtot, _ = self._wino_geometry(B, H, W)
        v_hi, v_lo = pool.get((36, mtot, cin), torch.float16), pool.get((36, mtot, cin), torch.float16)
        fkw = {} if film is None else dict(film_scale=film[0], film_shift=film[1], film_stride=film[2])
        rkw = {} if raw is None else dict(raw_hi=raw[0], raw_lo=raw[1])
        be.wino_input(src1, src2, groups=GN_GROUPS, mean=mean, rstd=rstd, gamma=gamma, beta=beta, silu=silu,
                      v_hi=v_hi, v_lo=v_lo, **fkw, **rkw)
        mbuf = pool.get((36, mtot, cout))
        be.conv_umma(B=36, H=mtot // 16, W=16, Cin=cin, Cout=cout, taps=1, a_hi=v_hi, a_lo=v_lo, w_hi=ent["u_hi"],
                     w_lo=ent["u_lo"], out=mbuf, passes=3, weights_per_image=True, operand_f16=True)
        pool.put(v_hi, v_lo)
        out = pool.get((B, H, W, cout))
        part = pool.get((B * th, cout, 2)) if stats else None
        be.wino_output(mbuf, B=B, H=H, W=W, Cout=cout, bias=ent["bias"], residual=residual, res_mode=res_mode,
                       out=out, stats_partial=part)
        pool.put(mbuf)
        if part is not None:
            out._gn = (part, th)
        return out


class UNetEngine(KernelExecutor):
    def __init__(self, unet: UNetModel, backend=None, precision: str = "split3"):
        super().__init__(backend, precision)
        self.unet = unet
        self._wkey = None
        self._w = {}
        self._table = None
        self.num_timesteps = 1000
        self.attention_impl = "tcgen05"      # "mma.sync" selects bbdm_attention_split for head_dim 64 too
        self.generation = 0          # bumps whenever cache/parameter ADDRESSES change (graphs key on it)

    # ------------------------------------------------------------------------------ weights
    def _params_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.unet.parameters())

    @staticmethod
    def _ptrs(key):
        return None if key is None else tuple(k[0] for k in key)

    def refresh_weights(self, force=False):
        """(Re)derive the packed weight caches if any parameter changed (optimizer step, EMA
        swap, load_state_dict).  The nn.Parameters themselves stay OIHW fp32."""
        key = self._params_key()
        if not force and key == self._wkey:
            return
        be, u = self.be, self.unet
        dev = next(u.parameters()).device
        # Derived caches are re-packed into the EXISTING buffers whenever name, shape and device match (no
        # reallocation of the ~5 GB of planes when the runner swaps EMA weights in and out for validation /
        # sampling, runners/base/EMA.py:31-43).  Same parameter storage (in-place optimizer update): every address a
        # captured CUDA graph holds stays valid; changed storage (EMA .data swap, load_state_dict): biases / norm
        # affines are read through the parameters' own addresses, so graphs are re-captured (generation bump).
        old = self._w or {}
        if self._ptrs(key) != self._ptrs(self._wkey) or not self._w:
            self.generation += 1
        w = {}

        def buf(name, field, shape, dtype):
            t = old.get(name, {}).get(field) if isinstance(old.get(name), dict) else None
            if t is not None and tuple(t.shape) == tuple(shape) and t.device == dev:
                return t
            return be.empty(tuple(shape), dtype, dev)

        def pack(conv, name):
            pack_tensor(conv.weight, conv.bias, name)

        def pack_tensor(weight, bias, name):
            wt = weight.detach()
            while wt.dim() < 4:                     # Conv1d [Cout, Cin, 1], Linear [out, in]
                wt = wt.unsqueeze(-1)
            wt = wt.contiguous()
            cout, cin, k = wt.shape[0], wt.shape[1], wt.shape[2]
            ent = {"cout": cout, "cin": cin, "k": k, "bias": bias.detach() if bias is not None else None}
            if cin % 64 == 0 and cout % 64 == 0 and k in (1, 3):
                hi = buf(name, "hi", (k * k, cout, cin), torch.bfloat16)
                lo = buf(name, "lo", (k * k, cout, cin), torch.bfloat16)
                be.pack_weight_split(wt, hi, lo)
                ent["hi"], ent["lo"] = hi, lo
            elif name == "out.2" and cin % 64 == 0 and cout < 64 and k == 3:
                # UNet head (Cout = 3..16): zero-padded to one 64-wide N tile of the tensor-core conv
                prev = old[name].get("hi_pad") if isinstance(old.get(name), dict) else None
                hi = buf(name, "hi_pad", (k * k, 64, cin), torch.bfloat16)
                lo = buf(name, "lo_pad", (k * k, 64, cin), torch.bfloat16)
                made = hi is not prev                 # freshly allocated: the padding rows must be zeroed
                bp = buf(name, "bias_pad", (64,), torch.float32)
                if made:
                    hi.zero_(); lo.zero_()
                bp.zero_()
                if bias is not None:
                    bp[:cout].copy_(bias.detach())
                be.pack_weight_split(wt, hi, lo)
                ent["hi_pad"], ent["lo_pad"], ent["bias_pad"] = hi, lo, bp
            f32 = buf(name, "f32", (k * k, cin, cout), torch.float32)
            be.pack_weight_f32(wt, f32)
            ent["f32"] = f32
            w[name] = ent

        film_w, film_b, off = [], [], 0
        for name, m in u.named_modules():
            if isinstance(m, (nn.Conv2d, nn.Conv1d)):
                pack(m, name)
            if isinstance(m, ResBlock):
                lin = m.emb_layers[1]
                n = lin.weight.shape[0]
                assert n % 4 == 0
                b = lin.bias.detach()
                if not m.use_scale_shift_norm:
                    # h + emb_out happens before out_layers' GroupNorm: fold conv1's bias into the
                    # per-sample vector so conv1 can take it as its (per-sample) bias directly
                    b = b + m.in_layers[2].bias.detach()
                film_w.append(lin.weight.detach())
                film_b.append(b)
                w[name + "#film"] = (off, n)
                off += n
        for name, m in u.named_modules():
            if isinstance(m, SpatialTransformer):
                # every nn.Linear of the transformer blocks as a 1x1 "conv" over the token grid; q|k|v of the
                # self-attention (and k|v of the cross-attention) concatenated into one GEMM each
                for j, blk in enumerate(m.transformer_blocks):
                    pre = f"{name}.transformer_blocks.{j}"
                    a1, a2 = blk.attn1, blk.attn2
                    pack_tensor(torch.cat([a1.to_q.weight, a1.to_k.weight, a1.to_v.weight], 0), None, pre + ".attn1.qkv")
                    pack_tensor(a1.to_out[0].weight, a1.to_out[0].bias, pre + ".attn1.to_out.0")
                    pack_tensor(a2.to_q.weight, None, pre + ".attn2.to_q")
                    pack_tensor(torch.cat([a2.to_k.weight, a2.to_v.weight], 0), None, pre + ".attn2.to_kv")
                    pack_tensor(a2.to_out[0].weight, a2.to_out[0].bias, pre + ".attn2.to_out.0")
                    pack_tensor(blk.ff.net[0].proj.weight, blk.ff.net[0].proj.bias, pre + ".ff.net.0.proj")
                    pack_tensor(blk.ff.net[2].weight, blk.ff.net[2].bias, pre + ".ff.net.2")
        if self.wino:
            # stride-1 3x3 ResBlock convs: Winograd-domain weight planes U = 2^8 G g G^T, fp16 hi/lo [36][Cout][Cin]
            for name, m in u.named_modules():
                if not isinstance(m, ResBlock) or not m.use_scale_shift_norm:
                    continue
                for cname, conv, skip in ((name + ".in_layers.2", m.in_layers[2], m.up or m.down),
                                          (name + ".out_layers.3", m.out_layers[3], False)):
                    ent = w[cname]
                    if skip or "hi" not in ent or ent["k"] != 3 or min(ent["cin"], ent["cout"]) < self.wino_min_c:
                        continue
                    uh = buf(cname, "u_hi", (36, ent["cout"], ent["cin"]), torch.float16)
                    ul = buf(cname, "u_lo", (36, ent["cout"], ent["cin"]), torch.float16)
                    be.wino_pack_weight(conv.weight.detach().contiguous(), uh, ul)
                    ent["u_hi"], ent["u_lo"] = uh, ul
        for name, m in u.named_modules():
            if isinstance(m, ResBlock) and m.up and m.channels % 64 == 0 and m.out_channels % 64 == 0:
                # up-ResBlock in_layers conv: 16 phase taps of the fused nearest-2x + 3x3 conv
                cname = name + ".in_layers.2"
                wp = upsample_phase_weights(m.in_layers[2].weight.detach())
                ph = buf(cname, "up_hi", (16, m.out_channels, m.channels), torch.bfloat16)
                pl = buf(cname, "up_lo", (16, m.out_channels, m.channels), torch.bfloat16)
                be.pack_weight_split_taps(wp, ph, pl)
                w[cname]["up_hi"], w[cname]["up_lo"] = ph, pl
        if old and old.get("film_n") == off and old["film_w"].device == dev:
            w["film_w"], w["film_b"] = old["film_w"], old["film_b"]
            torch.cat(film_w, 0, out=w["film_w"])
            torch.cat(film_b, 0, out=w["film_b"])
        else:
            w["film_w"] = torch.cat(film_w, 0).contiguous()
            w["film_b"] = torch.cat(film_b, 0).contiguous()
        w["film_n"] = off
        self._w = w
        self._wkey = key

    def _embedding_table(self, dev):
        """Rows 0..T-1 of the sinusoidal embedding (host-built with the reference's own expression,
        util.py:151-171: indexing is exact).  T follows the owning bridge model (``unet.num_timesteps``, set by
        BrownianBridgeModel.__init__) and is re-checked on every forward, outside the weight-key early return; an
        index >= T sets the device fault word in bbdm_gather_rows (the reference computes the embedding for any
        t, but its schedule gather raises for t >= T long before)."""
        want = max(self.num_timesteps, int(getattr(self.unet, "num_timesteps", 0) or 0))
        if self._table is None or self._table.shape[0] < want or self._table.device != dev:
            self.num_timesteps = want
            if self._table is not None:
                self.generation += 1      # table address changes: captured graphs must be rebuilt
            tab = timestep_embedding(torch.arange(want), self.unet.model_channels)
            self._table = tab.to(dev).contiguous()
        return self._table

    # ------------------------------------------------------------------------------ blocks
    def _resblock(self, pool, name, m: ResBlock, src1, src2, film):
        be, w = self.be, self._w
        B, Hs, Ws, c1 = src1.shape
        c2 = 0 if src2 is None else src2.shape[3]
        cin, cout = c1 + c2, m.out_channels
        assert cin == m.channels
        resample = cabi.RESAMPLE_UP2 if m.up else (cabi.RESAMPLE_DOWN2 if m.down else cabi.RESAMPLE_NONE)
        H, W = (Hs * 2, Ws * 2) if m.up else ((Hs // 2, Ws // 2) if m.down else (Hs, Ws))
        e1, e2 = w[name + ".in_layers.2"], w[name + ".out_layers.3"]
        skip_conv = isinstance(m.skip_connection, nn.Conv2d)
        es = w[name + ".skip_connection"] if skip_conv else None
        umma1 = self._umma_ok(cin, cout, W)
        umma2 = self._umma_ok(cout, cout, W)
        fuse_skip = skip_conv and umma2 and es["k"] == 1 and cin % 64 == 0
        need_raw_f32 = (skip_conv and not fuse_skip) or \
                       (not skip_conv and (src2 is not None or (resample != cabi.RESAMPLE_NONE and not umma2)))
        foff, fn = w[name + "#film"]

        # ---- in_layers: GN -> SiLU -> (up/down) -> conv3x3 -------------------------------------
        mean, rstd = self._stats(pool, src1, src2)
        gn = m.in_layers[0]
        # up-ResBlock on the tensor-core path: never materialise the upsampled activation -- the
        # conv runs as 4 output phases x 2x2 taps on the low-res operand (2.25x fewer MACs)
        fused_up = bool(m.up and umma1 and "up_hi" in e1 and Ws >= 4 and m.use_scale_shift_norm
                        and not need_raw_f32 and not fuse_skip)
        if fused_up:
            a_hi, a_lo = pool.get((B, Hs, Ws, cin), torch.bfloat16), pool.get((B, Hs, Ws, cin), torch.bfloat16)
            be.prep(src1, src2, groups=GN_GROUPS, mean=mean, rstd=rstd, gamma=gn.weight.detach(),
                    beta=gn.bias.detach(), silu=True, resample=cabi.RESAMPLE_NONE, act_hi=a_hi, act_lo=a_lo)
            pool.put(mean, rstd)
            h1 = pool.get((B, H, W, cout))
            rows = 4 * self._geom(Hs, Ws)
            part = pool.get((B * rows, cout, 2)) if rows else None
            be.conv_umma(B=B, H=Hs, W=Ws, Cin=cin, Cout=cout, taps=4, a_hi=a_hi, a_lo=a_lo, w_hi=e1["up_hi"],
                         w_lo=e1["up_lo"], bias=e1["bias"], out=h1, passes=self.passes, upsample2x=True,
                         stats_partial=part)
            if part is not None:
                h1._gn = (part, rows)
            pool.put(a_hi, a_lo)
            return self._resblock_tail(pool, name, m, src1, h1, film, foff, cout, (B, H, W), e2, es, umma2,
                                       None, cabi.RES_UP2, None, None, None)
        shp = (B, H, W, cin)
        a_f32 = a_hi = a_lo = r_f32 = r_hi = r_lo = None
        if umma1 and resample == cabi.RESAMPLE_NONE and not need_raw_f32 and m.use_scale_shift_norm \
                and self._wino_ok(e1, B, H, W):
            # Winograd conv1; the raw split planes for a fused 1x1 skip come out of the same input pass
            if fuse_skip:
                r_hi, r_lo = pool.get(shp, torch.bfloat16), pool.get(shp, torch.bfloat16)
            h1 = self._wino_conv(pool, e1, src1, src2, mean=mean, rstd=rstd, gamma=gn.weight.detach(),
                                 beta=gn.bias.detach(), raw=(r_hi, r_lo) if fuse_skip else None)
            pool.put(mean, rstd)
            return self._resblock_tail(pool, name, m, src1, h1, film, foff, cout, (B, H, W), e2, es, umma2,
                                       (es, r_hi, r_lo) if fuse_skip else None, resample_to_res(resample),
                                       None, r_hi, r_lo, skip_conv=skip_conv, need_raw_f32=False)
        if umma1:
            a_hi, a_lo = pool.get(shp, torch.bfloat16), pool.get(shp, torch.bfloat16)
        else:
            a_f32 = pool.get(shp)
        if fuse_skip:
            r_hi, r_lo = pool.get(shp, torch.bfloat16), pool.get(shp, torch.bfloat16)
        if need_raw_f32:
            r_f32 = pool.get(shp)
        be.prep(src1, src2, groups=GN_GROUPS, mean=mean, rstd=rstd, gamma=gn.weight.detach(), beta=gn.bias.detach(),
                silu=True, resample=resample, act_f32=a_f32, act_hi=a_hi, act_lo=a_lo,
                raw_f32=r_f32, raw_hi=r_hi, raw_lo=r_lo)
        pool.put(mean, rstd)
        if m.use_scale_shift_norm:
            h1, _, _ = self._conv(pool, e1, a_f32=a_f32, a_hi=a_hi, a_lo=a_lo, shape=(B, H, W), stats=True)
        else:
            # conv1 + (bias + emb_out[b]) per sample: per-sample bias rows live in `film`
            h1 = pool.get((B, H, W, cout))
            for b in range(B):
                sl = lambda z: None if z is None else z[b:b + 1]
                self._conv(pool, e1, a_f32=sl(a_f32), a_hi=sl(a_hi), a_lo=sl(a_lo), shape=(1, H, W),
                           bias=film[b, foff:foff + fn], out=h1[b:b + 1])
        pool.put(a_f32, a_hi, a_lo)

        if fuse_skip:
            second_args = (es, r_hi, r_lo)
        else:
            second_args = None
        return self._resblock_tail(pool, name, m, src1, h1, film, foff, cout, (B, H, W), e2, es, umma2,
                                   second_args, resample_to_res(resample), r_f32, r_hi, r_lo,
                                   skip_conv=skip_conv, need_raw_f32=need_raw_f32)

    def _resblock_tail(self, pool, name, m, src1, h1, film, foff, cout, shape, e2, es, umma2, second, id_res_mode,
                       r_f32, r_hi, r_lo, skip_conv=False, need_raw_f32=False):
        """out_layers of a ResBlock: GN (+FiLM) -> SiLU -> conv3x3 with the skip path fused in."""
        be = self.be
        B, H, W = shape
        # ---- out_layers: GN (+FiLM) -> SiLU -> conv3x3 (+skip) -----------------------------------
        mean, rstd = self._stats(pool, h1, None)
        gn2 = m.out_layers[0]
        shp2 = (B, H, W, cout)
        b_f32 = b_hi = b_lo = None
        if umma2 and m.use_scale_shift_norm and self._wino_ok(e2, B, H, W):
            # Winograd conv2: the 1x1 skip (if any) runs as its own tensor-core GEMM and enters as the residual
            residual, res_mode, skip_out = None, cabi.RES_NONE, None
            if second is not None:
                skip_out, _, _ = self._conv(pool, second[0], a_hi=second[1], a_lo=second[2], shape=(B, H, W))
                residual, res_mode = skip_out, cabi.RES_SAME
            elif skip_conv:
                skip_out, _, _ = self._conv(pool, es, a_f32=r_f32, shape=(B, H, W))
                residual, res_mode = skip_out, cabi.RES_SAME
            elif need_raw_f32:
                residual, res_mode = r_f32, cabi.RES_SAME
            else:
                residual, res_mode = src1, id_res_mode
            out = self._wino_conv(pool, e2, h1, None, mean=mean, rstd=rstd, gamma=gn2.weight.detach(),
                                  beta=gn2.bias.detach(),
                                  film=(film[:, foff:foff + cout], film[:, foff + cout:foff + 2 * cout], film.shape[1]),
                                  residual=residual, res_mode=res_mode)
            pool.put(mean, rstd, h1, r_f32, r_hi, r_lo, skip_out)
            return out
        if umma2:
            b_hi, b_lo = pool.get(shp2, torch.bfloat16), pool.get(shp2, torch.bfloat16)
        else:
            b_f32 = pool.get(shp2)
        fkw = {}
        if m.use_scale_shift_norm:
            fkw = dict(film_scale=film[:, foff:foff + cout], film_shift=film[:, foff + cout:foff + 2 * cout],
                       film_stride=film.shape[1])
        be.prep(h1, None, groups=GN_GROUPS, mean=mean, rstd=rstd, gamma=gn2.weight.detach(), beta=gn2.bias.detach(),
                silu=True, resample=cabi.RESAMPLE_NONE, act_f32=b_f32, act_hi=b_hi, act_lo=b_lo, **fkw)
        pool.put(mean, rstd, h1)

        residual, res_mode, skip_out = None, cabi.RES_NONE, None
        if second is not None:
            pass
        elif skip_conv:
            skip_out, _, _ = self._conv(pool, es, a_f32=r_f32, shape=(B, H, W))
            residual, res_mode = skip_out, cabi.RES_SAME
        elif need_raw_f32:
            residual, res_mode = r_f32, cabi.RES_SAME
        else:
            residual, res_mode = src1, id_res_mode
        out, _, _ = self._conv(pool, e2, a_f32=b_f32, a_hi=b_hi, a_lo=b_lo, shape=(B, H, W),
                               residual=residual, res_mode=res_mode, second=second, stats=True)
        pool.put(b_f32, b_hi, b_lo, r_f32, r_hi, r_lo, skip_out)
        return out

    def _attention(self, pool, name, m: AttentionBlock, x):
        be, w = self.be, self._w
        B, H, W, Cc = x.shape
        T = H * W
        eq, ep = w[name + ".qkv"], w[name + ".proj_out"]
        heads = m.num_heads
        hd = Cc // heads
        if hd not in (16, 32, 64):
            raise NotImplementedError(f"attention head_dim {hd}: the sm_100a kernel supports 16/32/64")
        umma = self._umma_ok(Cc, Cc, W)
        mean, rstd = self._stats(pool, x, None)
        a_f32 = a_hi = a_lo = None
        if umma:
            a_hi, a_lo = pool.get(x.shape, torch.bfloat16), pool.get(x.shape, torch.bfloat16)
        else:
            a_f32 = pool.get(x.shape)
        be.prep(x, None, groups=GN_GROUPS, mean=mean, rstd=rstd, gamma=m.norm.weight.detach(),
                beta=m.norm.bias.detach(), silu=False, resample=cabi.RESAMPLE_NONE,
                act_f32=a_f32, act_hi=a_hi, act_lo=a_lo)
        pool.put(mean, rstd)
        # qkv 1x1: on the tensor-core path its epilogue writes the split planes the attention core reads
        qkv, q_hi, q_lo = self._conv(pool, eq, a_f32=a_f32, a_hi=a_hi, a_lo=a_lo, shape=(B, H, W),
                                     out_split=umma, want_f32=not umma)
        pool.put(a_f32, a_hi, a_lo)
        o_f32 = o_hi = o_lo = None
        order = 1 if m.new_order else 0
        if umma:
            o_hi, o_lo = pool.get(x.shape, torch.bfloat16), pool.get(x.shape, torch.bfloat16)
            # head_dim 64 (all templates): warp-specialised tcgen05 kernel; else the mma.sync one
            attn = be.attention_tc if (hd == 64 and self.attention_impl == "tcgen05") else be.attention_split
            attn(q_hi.view(B, T, 3 * Cc), q_lo.view(B, T, 3 * Cc), heads, order,
                 None, o_hi.view(B, T, Cc), o_lo.view(B, T, Cc))
        else:
            o_f32 = pool.get(x.shape)
            be.attention(qkv.view(B, T, 3 * Cc), heads, order, o_f32.view(B, T, Cc), None, None)
        pool.put(qkv, q_hi, q_lo)
        out, _, _ = self._conv(pool, ep, a_f32=o_f32, a_hi=o_hi, a_lo=o_lo, shape=(B, H, W),
                               residual=x, res_mode=cabi.RES_SAME, stats=True)
        pool.put(o_f32, o_hi, o_lo)
        return out

    def _stem(self, pool, ent, x):
        """A bare nn.Conv2d inside a block = the UNet stem (openaimodel.py:524): few input channels.  Dedicated kernel
        (weights in shared memory, GroupNorm partial sums of the output fused) when the shape fits, else the general
        fp32 kernel."""
        B, H, W, cin = x.shape
        cout = ent["cout"]
        if ent["k"] == 3 and cin <= 16 and cout % 32 == 0 and cout <= 128 and W % 32 == 0 and hasattr(self.be, "conv_stem"):
            out = pool.get((B, H, W, cout))
            part = pool.get((B * H, cout, 2))
            self.be.conv_stem(x, ent["f32"], ent["bias"], out, cout, stats_partial=part)
            out._gn = (part, H)
            return out
        out, _, _ = self._conv(pool, ent, a_f32=x, shape=(B, H, W))
        return out

    def _spatial_transformer(self, pool, name, m: SpatialTransformer, x, ctx):
        """GroupNorm(1e-6) -> proj_in -> [LN -> self-attention -> +, LN -> cross-attention(context) -> +,
        LN -> GEGLU feed-forward -> +] x depth -> proj_out -> + x   (reference attention.py:196-264).  Every Linear /
        1x1 conv is a tcgen05 GEMM over the token grid whose epilogue adds the residual; LayerNorm and GEGLU write the
        next GEMM's split operand planes directly."""
        be, w = self.be, self._w
        B, H, W, Cc = x.shape
        T, heads, d = H * W, m.n_heads, m.d_head
        inner = heads * d
        if d not in (16, 32, 64):
            raise NotImplementedError(f"SpatialTransformer head_dim {d}: the sm_100a attention kernels take 16/32/64")
        if not (self._umma_ok(Cc, inner, W) and inner % 64 == 0):
            raise NotImplementedError("SpatialTransformer: channel counts must be multiples of 64 (tensor-core GEMMs)")
        bf = torch.bfloat16
        tok = (B, H, W, inner)
        mean, rstd = self._stats(pool, x, None, eps=m.norm.eps)
        a_hi, a_lo = pool.get(x.shape, bf), pool.get(x.shape, bf)
        be.prep(x, None, groups=GN_GROUPS, mean=mean, rstd=rstd, gamma=m.norm.weight.detach(), beta=m.norm.bias.detach(),
                silu=False, resample=cabi.RESAMPLE_NONE, act_hi=a_hi, act_lo=a_lo)
        pool.put(mean, rstd)
        h, _, _ = self._conv(pool, w[name + ".proj_in"], a_hi=a_hi, a_lo=a_lo, shape=(B, H, W))
        pool.put(a_hi, a_lo)

        def layernorm(ln, src):
            n_hi, n_lo = pool.get(tok, bf), pool.get(tok, bf)
            be.layernorm_split(src, ln.weight.detach(), ln.bias.detach(), ln.eps, out_hi=n_hi, out_lo=n_lo)
            return n_hi, n_lo

        def project_out(ent, o_hi, o_lo, res):
            new, _, _ = self._conv(pool, ent, a_hi=o_hi, a_lo=o_lo, shape=(B, H, W), residual=res, res_mode=cabi.RES_SAME)
            pool.put(o_hi, o_lo, res)
            return new

        for j, blk in enumerate(m.transformer_blocks):
            pre = f"{name}.transformer_blocks.{j}"
            # ---- self-attention ----------------------------------------------------------------------------------
            n_hi, n_lo = layernorm(blk.norm1, h)
            _, q_hi, q_lo = self._conv(pool, w[pre + ".attn1.qkv"], a_hi=n_hi, a_lo=n_lo, shape=(B, H, W), out_split=True,
                                       want_f32=False)
            pool.put(n_hi, n_lo)
            o_hi, o_lo = pool.get(tok, bf), pool.get(tok, bf)
            attn = be.attention_tc if (d == 64 and self.attention_impl == "tcgen05") else be.attention_split
            attn(q_hi.view(B, T, 3 * inner), q_lo.view(B, T, 3 * inner), heads, 1, None, o_hi.view(B, T, inner),
                 o_lo.view(B, T, inner))
            pool.put(q_hi, q_lo)
            h = project_out(w[pre + ".attn1.to_out.0"], o_hi, o_lo, h)
            # ---- cross-attention over the conditioning tokens (or over h itself without a context) ------------------
            n_hi, n_lo = layernorm(blk.norm2, h)
            _, q_hi, q_lo = self._conv(pool, w[pre + ".attn2.to_q"], a_hi=n_hi, a_lo=n_lo, shape=(B, H, W), out_split=True,
                                       want_f32=False)
            ekv = w[pre + ".attn2.to_kv"]
            if ctx is None:
                _, kv_hi, kv_lo = self._conv(pool, ekv, a_hi=n_hi, a_lo=n_lo, shape=(B, H, W), out_split=True, want_f32=False)
                Tc = T
            else:
                Bc, Hc, Wc, _ = ctx.shape
                Tc = Hc * Wc
                kv, _, _ = self._conv(pool, ekv, a_f32=ctx, shape=(Bc, Hc, Wc))          # few context channels: fp32 direct
                kv_hi, kv_lo = pool.get(kv.shape, bf), pool.get(kv.shape, bf)
                be.prep(kv, None, resample=cabi.RESAMPLE_NONE, raw_hi=kv_hi, raw_lo=kv_lo)
                pool.put(kv)
            pool.put(n_hi, n_lo)
            o_hi, o_lo = pool.get(tok, bf), pool.get(tok, bf)
            be.attention_cross(q_hi.view(B, T, inner), q_lo.view(B, T, inner), kv_hi.view(B, Tc, 2 * inner),
                               kv_lo.view(B, Tc, 2 * inner), heads, None, o_hi.view(B, T, inner), o_lo.view(B, T, inner))
            pool.put(q_hi, q_lo, kv_hi, kv_lo)
            h = project_out(w[pre + ".attn2.to_out.0"], o_hi, o_lo, h)
            # ---- GEGLU feed-forward -----------------------------------------------------------------------------------
            if not blk.ff.glu:
                raise NotImplementedError("SpatialTransformer feed-forward without GEGLU")
            n_hi, n_lo = layernorm(blk.norm3, h)
            uu, _, _ = self._conv(pool, w[pre + ".ff.net.0.proj"], a_hi=n_hi, a_lo=n_lo, shape=(B, H, W))
            pool.put(n_hi, n_lo)
            ffi = uu.shape[3] // 2
            g_hi, g_lo = pool.get((B, H, W, ffi), bf), pool.get((B, H, W, ffi), bf)
            be.geglu_split(uu, out_hi=g_hi, out_lo=g_lo)
            pool.put(uu)
            h = project_out(w[pre + ".ff.net.2"], g_hi, g_lo, h)
        r_hi, r_lo = pool.get(tok, bf), pool.get(tok, bf)
        be.prep(h, None, resample=cabi.RESAMPLE_NONE, raw_hi=r_hi, raw_lo=r_lo)
        pool.put(h)
        out, _, _ = self._conv(pool, w[name + ".proj_out"], a_hi=r_hi, a_lo=r_lo, shape=(B, H, W), residual=x,
                               res_mode=cabi.RES_SAME, stats=True)
        pool.put(r_hi, r_lo)
        return out

    def _resample_layer(self, pool, name, m, x):
        be, w = self.be, self._w
        B, H, W, Cc = x.shape
        if isinstance(m, Downsample):
            if m.use_conv:
                out, _, _ = self._conv(pool, w[name + ".op"], a_f32=x, shape=(B, H, W), stride=2)
                return out
            out = pool.get((B, H // 2, W // 2, Cc))
            be.prep(x, None, resample=cabi.RESAMPLE_DOWN2, raw_f32=out)
            return out
        up = pool.get((B, H * 2, W * 2, Cc))
        be.prep(x, None, resample=cabi.RESAMPLE_UP2, raw_f32=up)
        if not m.use_conv:
            return up
        out, _, _ = self._conv(pool, w[name + ".conv"], a_f32=up, shape=(B, H * 2, W * 2))
        pool.put(up)
        return out

    def _run_block(self, pool, prefix, block: TimestepEmbedSequential, h, skip, film, release_input):
        """h (+ skip, channel-concatenated behind it for the first layer) through one block.
        release_input: whether h/skip may go back to the pool once the first layer has consumed
        them (False while they are still referenced as saved skip tensors)."""
        cur, cur_skip, owned = h, skip, release_input
        for j, layer in enumerate(block):
            name = f"{prefix}.{j}"
            if isinstance(layer, ResBlock):
                new = self._resblock(pool, name, layer, cur, cur_skip, film)
            else:
                assert cur_skip is None, "a concatenated input is only consumed by a ResBlock"
                if isinstance(layer, AttentionBlock):
                    new = self._attention(pool, name, layer, cur)
                elif isinstance(layer, SpatialTransformer):
                    new = self._spatial_transformer(pool, name, layer, cur, self._ctx_nhwc)
                elif isinstance(layer, (Downsample, Upsample)):
                    new = self._resample_layer(pool, name, layer, cur)
                elif isinstance(layer, nn.Conv2d):
                    new = self._stem(pool, self._w[name], cur)
                else:
                    raise NotImplementedError(type(layer).__name__)
            if owned:
                pool.put(cur, cur_skip)
            cur, cur_skip, owned = new, None, True
        return cur

    # ------------------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x, timesteps, context=None, assume_fresh_weights=False, out=None):
        u, be = self.unet, self.be
        if not assume_fresh_weights:
            self.refresh_weights()
        w = self._w
        dev = x.device
        x = x.contiguous().float()
        B, Cx, H, W = x.shape
        pool = self._pool(dev, (B, H, W))
        ctx = None
        if u.condition_key != "nocond":
            ctx = context.contiguous().float()
        t = timesteps.to(device=dev, dtype=torch.int64).contiguous()

        # ---- timestep embedding MLP + all FiLM projections (3 small fp32 GEMV launches) --------
        mc, ted = u.model_channels, u.model_channels * 4
        temb, e1, emb = pool.get((B, mc)), pool.get((B, ted)), pool.get((B, ted))
        film = pool.get((B, w["film_n"]))
        be.gather_rows(self._embedding_table(dev), t, temb)
        l0, l2 = u.time_embed[0], u.time_embed[2]
        be.linear(temb, l0.weight.detach(), l0.bias.detach(), e1, act_out=True)
        be.linear(e1, l2.weight.detach(), l2.bias.detach(), emb)
        be.linear(emb, w["film_w"], w["film_b"], film, act_in=True)
        pool.put(temb, e1)

        # ---- stem --------------------------------------------------------------------------------
        cin0 = Cx + (0 if ctx is None else ctx.shape[1])
        xin = pool.get((B, H, W, cin0))
        be.nchw_to_nhwc_cat(x, ctx, xin)
        self._ctx_nhwc = None
        if getattr(u, "use_spatial_transformer", False) and ctx is not None:
            # the transformers cross-attend to the same conditioning tensor (openaimodel.py:745-748), token-major
            self._ctx_nhwc = pool.get((B, ctx.shape[2], ctx.shape[3], ctx.shape[1]))
            be.nchw_to_nhwc_cat(ctx, None, self._ctx_nhwc)
        hs = []
        h = xin
        for i, block in enumerate(u.input_blocks):
            # block inputs after the stem are saved skip tensors (hs): not released here
            h = self._run_block(pool, f"input_blocks.{i}", block, h, None, film, release_input=(i == 0))
            hs.append(h)
        h = self._run_block(pool, "middle_block", u.middle_block, h, None, film, release_input=False)
        for i, block in enumerate(u.output_blocks):
            # h is the previous block's output, hs.pop() the matching skip: both die here
            h = self._run_block(pool, f"output_blocks.{i}", block, h, hs.pop(), film, release_input=True)

        # ---- head: GN -> SiLU -> conv3x3 -> NCHW ----------------------------------------------------
        mean, rstd = self._stats(pool, h, None)
        gn = u.out[0]
        if out is None:
            out = torch.empty((B, u.out_channels, H, W), dtype=torch.float32, device=dev)
        eh = w["out.2"]
        if "hi_pad" in eh and W >= 4:
            # tensor-core head: N tile padded to 64 couts, epilogue stores the real ones as NCHW
            a_hi, a_lo = pool.get(h.shape, torch.bfloat16), pool.get(h.shape, torch.bfloat16)
            be.prep(h, None, groups=GN_GROUPS, mean=mean, rstd=rstd, gamma=gn.weight.detach(), beta=gn.bias.detach(),
                    silu=True, resample=cabi.RESAMPLE_NONE, act_hi=a_hi, act_lo=a_lo)
            pool.put(mean, rstd, h)
            be.conv_umma(B=B, H=H, W=W, Cin=eh["cin"], Cout=64, taps=9, a_hi=a_hi, a_lo=a_lo, w_hi=eh["hi_pad"],
                         w_lo=eh["lo_pad"], bias=eh["bias_pad"], out=out, passes=self.passes,
                         out_nchw_channels=u.out_channels)
            pool.put(a_hi, a_lo, emb, film, self._ctx_nhwc)
            self._ctx_nhwc = None
            return out
        act = pool.get(h.shape)
        be.prep(h, None, groups=GN_GROUPS, mean=mean, rstd=rstd, gamma=gn.weight.detach(), beta=gn.bias.detach(),
                silu=True, resample=cabi.RESAMPLE_NONE, act_f32=act)
        pool.put(mean, rstd, h)
        y, _, _ = self._conv(pool, eh, a_f32=act, shape=(B, H, W))
        pool.put(act)
        be.nhwc_to_nchw(y, out)
        pool.put(y, emb, film, self._ctx_nhwc)
        self._ctx_nhwc = None
        return out
