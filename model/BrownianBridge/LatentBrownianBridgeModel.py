"""Drop-in ``LatentBrownianBridgeModel``: the latent-space wrapper around the B200-native
bridge model.  Same surface as the reference class
(/root/reference/model/BrownianBridge/LatentBrownianBridgeModel.py:19-132): frozen VQGAN
encode/decode at both ends (kept as the reference's own PyTorch ``VQModel`` per the north star -- as the parameter container; on a GPU its encoder / quantiser / decoder run
through ``bbdm_b200.vqgan_engine.VQGANEngine`` on the sm_100a kernels),
optional cond-stage model, ``encode`` / ``decode`` / ``sample`` / ``sample_vqgan`` /
``get_ema_net`` and the latent mean/std attributes the runner assigns
(BBDMRunner.py:41-44,136-158).  Everything between encode and decode -- q_sample, the UNet,
the p_sample loop -- runs on the sm_100a kernels via the base class.
"""
import itertools
import os
import warnings

import torch

from bbdm_b200.cond import SpatialRescaler
from model.BrownianBridge.BrownianBridgeModel import BrownianBridgeModel

try:
    from tqdm.autonotebook import tqdm
except Exception:  # pragma: no cover
    def tqdm(it, **kw):
        return it


def _frozen_train(self, mode=True):
    """train()/eval() become no-ops on the frozen autoencoder."""
    return self


def _load_vqmodel():
    try:
        from model.VQGAN.vqgan import VQModel   # the reference's module, unmodified
    except ImportError:
        # no BBDM checkout on sys.path: the repo's own parameter tree (same state_dict names); GPU only
        from bbdm_b200.vqgan import VQModel
    return VQModel


class LatentBrownianBridgeModel(BrownianBridgeModel):
    # frozen autoencoder on the sm_100a kernels when the tensors live on a GPU (BBDM_NATIVE_VQGAN=0: keep the
    # PyTorch module path)
    native_vqgan = os.environ.get("BBDM_NATIVE_VQGAN", "1") != "0"

    def __init__(self, model_config):
        super().__init__(model_config)
        self.vqgan = _load_vqmodel()(**vars(model_config.VQGAN.params)).eval()
        self.vqgan.train = _frozen_train
        for param in self.vqgan.parameters():
            param.requires_grad = False
        print(f"load vqgan from {model_config.VQGAN.params.ckpt_path}")

        if self.condition_key == 'nocond':
            self.cond_stage_model = None
        elif self.condition_key == 'first_stage':
            self.cond_stage_model = self.vqgan
        elif self.condition_key == 'SpatialRescaler':
            self.cond_stage_model = SpatialRescaler(**vars(model_config.CondStageParams))
        else:
            raise NotImplementedError

    def get_ema_net(self):
        return self

    def get_parameters(self):
        if self.condition_key == 'SpatialRescaler':
            print("get parameters to optimize: SpatialRescaler, UNet")
            return itertools.chain(self.denoise_fn.parameters(), self.cond_stage_model.parameters())
        print("get parameters to optimize: UNet")
        return self.denoise_fn.parameters()

    def apply(self, weights_init):
        super().apply(weights_init)
        if self.cond_stage_model is not None:
            # NB: with condition_key 'first_stage' this re-initialises the VQGAN, exactly
            # as the reference does (:51-55).
            self.cond_stage_model.apply(weights_init)
        return self

    # ---- training: both encodes are no-grad, only the UNet (+ rescaler) trains -----------------------
    def forward(self, x, x_cond, context=None):
        with torch.no_grad():
            x_latent = self.encode(x, cond=False)
            x_cond_latent = self.encode(x_cond, cond=True)
        context = self.get_cond_stage_context(x_cond)
        return super().forward(x_latent.detach(), x_cond_latent.detach(), context)

    def get_cond_stage_context(self, x_cond):
        if self.cond_stage_model is None:
            return None
        context = self.cond_stage_model(x_cond)
        return context.detach() if self.condition_key == 'first_stage' else context

    # ---- frozen autoencoder ends ------------------------------------------------------------------------
    def _norm_stats(self, cond):
        if cond:
            return self.cond_latent_mean, self.cond_latent_std
        return self.ori_latent_mean, self.ori_latent_std

    def _vq_engine(self):
        eng = self.__dict__.get("_vq_eng")
        if eng is None:
            from bbdm_b200.vqgan_engine import VQGANEngine
            eng = self.__dict__["_vq_eng"] = VQGANEngine(self.vqgan)
        return eng

    def _native(self, fn, x, *args):
        """Run the autoencoder end on the kernels; shapes they do not take fall back to the module (once warned)."""
        if not (self.native_vqgan and x.is_cuda):
            return None
        try:
            return fn(x, *args)
        except NotImplementedError as e:
            if not hasattr(self.vqgan.encoder, "forward") or type(self.vqgan).__module__.startswith("bbdm_b200"):
                raise
            warnings.warn(f"VQGAN on the PyTorch module path: {e}")
            self.native_vqgan = False
            return None

    @torch.no_grad()
    def encode(self, x, cond=True, normalize=None):
        normalize = self.model_config.normalize_latent if normalize is None else normalize
        before = self.model_config.latent_before_quant_conv
        z = self._native(lambda t: self._vq_engine().encode(t, quant_conv=not before), x)
        if z is None:
            z = self.vqgan.encoder(x)
            if not before:
                z = self.vqgan.quant_conv(z)
        if normalize:
            mean, std = self._norm_stats(cond)
            z = (z - mean) / std
        return z

    @torch.no_grad()
    def decode(self, x_latent, cond=True, normalize=None):
        normalize = self.model_config.normalize_latent if normalize is None else normalize
        if normalize:
            mean, std = self._norm_stats(cond)
            x_latent = x_latent * std + mean
        before = self.model_config.latent_before_quant_conv
        out = self._native(lambda t: self._vq_engine().decode(t, quant_conv_first=before), x_latent)
        if out is not None:
            return out
        if before:
            x_latent = self.vqgan.quant_conv(x_latent)
        x_latent_quant, _, _ = self.vqgan.quantize(x_latent)
        return self.vqgan.decode(x_latent_quant)

    @torch.no_grad()
    def sample(self, x_cond, clip_denoised=False, sample_mid_step=False):
        """NB: no ``context`` argument and clip_denoised defaults to False here (reference :102)."""
        x_cond_latent = self.encode(x_cond, cond=True)
        result = self.p_sample_loop(y=x_cond_latent, context=self.get_cond_stage_context(x_cond),
                                    clip_denoised=clip_denoised, sample_mid_step=sample_mid_step)
        if not sample_mid_step:
            return self.decode(result, cond=False)
        decoded = []
        for seq, desc in zip(result, ("save output sample mid steps", "save one step sample mid steps")):
            outs = []
            for z in tqdm(seq, initial=0, desc=desc, dynamic_ncols=True, smoothing=0.01):
                outs.append(self.decode(z.detach(), cond=False).to('cpu'))
            decoded.append(outs)
        return decoded[0], decoded[1]

    @torch.no_grad()
    def sample_vqgan(self, x):
        """Autoencoder round trip (reference :128-132: ``vqgan(x)`` = decode(quantize(quant_conv(encoder(x)))))."""
        out = self._native(lambda t: self._vq_engine().decode(self._vq_engine().encode(t, quant_conv=True)), x)
        if out is not None:
            return out
        x_rec, _ = self.vqgan(x)
        return x_rec
