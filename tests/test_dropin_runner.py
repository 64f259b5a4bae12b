"""Drop-in check: the UNMODIFIED reference runner (runners/DiffusionBasedModelRunners/BBDMRunner.py)
trains and samples on top of this repo's model classes, reached through the namespace-package
overlay.  CPU here (kernel backend emulated by the oracle -- tests only); needs /root/reference."""
import argparse
import os
import sys

import numpy as np
import pytest
import torch
import yaml

import conftest
from _emu_backend import EmuBackend

pytestmark = pytest.mark.reference


def _make_dataset(root, n=4, size=16):
    from PIL import Image
    rng = np.random.RandomState(0)
    for stage in ("train", "val", "test"):
        for side in ("A", "B"):
            d = os.path.join(root, stage, side)
            os.makedirs(d, exist_ok=True)
            for i in range(n):
                Image.fromarray(rng.randint(0, 255, (size, size, 3), dtype=np.uint8)).save(os.path.join(d, f"{i}.png"))


def test_unmodified_bbdm_runner_trains_and_samples(tmp_path, monkeypatch):
    conftest.install_reference_shims()
    from bbdm_b200.bridge import BridgeOps
    monkeypatch.setattr(BridgeOps, "backend_factory", staticmethod(lambda: EmuBackend()))
    from bbdm_b200.optim import FusedEMA
    monkeypatch.setattr(FusedEMA, "backend_factory", staticmethod(lambda: EmuBackend()))   # runners.base.EMA overlay
    # reference modules, untouched
    from utils import dict2namespace, get_runner
    import model.BrownianBridge.BrownianBridgeModel as overlay
    import runners.DiffusionBasedModelRunners.BBDMRunner as runner_mod
    assert overlay.__file__.startswith(conftest.REPO)                 # model classes: this repo
    assert runner_mod.__file__.startswith(conftest.REF)               # runner: the reference
    import runners.base.EMA as ema_mod
    assert ema_mod.__file__.startswith(conftest.REPO) and ema_mod.EMA is FusedEMA   # EMA: this repo's overlay
    assert runner_mod.BrownianBridgeModel is overlay.BrownianBridgeModel

    with open(os.path.join(conftest.REF, "configs", "Template-BBDM.yaml")) as f:
        cfg = yaml.load(f, Loader=yaml.FullLoader)
    data = str(tmp_path / "data")
    _make_dataset(data)
    cfg["data"]["dataset_config"].update(dataset_path=data, image_size=16)
    cfg["data"]["dataset_name"] = "synthetic"
    for s in ("train", "val", "test"):
        cfg["data"][s]["batch_size"] = 2
    cfg["training"].update(n_epochs=1, n_steps=2, save_interval=1, sample_interval=1, validation_interval=1,
                           accumulate_grad_batches=1)
    cfg["testing"]["sample_num"] = 1
    cfg["model"]["EMA"].update(start_ema_step=1, update_ema_interval=1)
    u = cfg["model"]["BB"]["params"]["UNetParams"]
    u.update(image_size=16, model_channels=32, num_head_channels=32)
    cfg["model"]["BB"]["params"]["sample_step"] = 4

    def run(train):
        nc = dict2namespace(cfg)
        nc.args = argparse.Namespace(config="x", seed=1234, result_path=str(tmp_path / "results"), train=train,
                                     sample_to_eval=not train, sample_at_start=False, save_top=False, gpu_ids="-1",
                                     port="12355", resume_model=None, resume_optim=None, max_epoch=None, max_steps=None)
        nc.training.use_DDP = False
        nc.training.device = [torch.device("cpu")]
        if not train:
            ck = os.path.join(str(tmp_path / "results"), "synthetic", "BrownianBridge", "checkpoint")
            nc.model.model_load_path = os.path.join(ck, "last_model.pth")
        torch.manual_seed(1234)
        runner = get_runner(nc.runner, nc)
        if train:
            runner.train()
        else:
            with torch.no_grad():
                runner.test()
        return runner

    monkeypatch.setattr(torch.utils.data.DataLoader, "__init__",
                        _no_workers(torch.utils.data.DataLoader.__init__))
    # torch >= 2.4 removed ReduceLROnPlateau(verbose=...), which the reference (torch 1.12) passes
    RLP = torch.optim.lr_scheduler.ReduceLROnPlateau
    monkeypatch.setattr(RLP, "__init__", _drop_kw(RLP.__init__, "verbose"))
    r = run(train=True)
    ck = os.path.join(r.config.result.ckpt_path, "last_model.pth")
    assert os.path.exists(ck)
    sd = torch.load(ck, map_location="cpu", weights_only=False)
    assert any(k.startswith("denoise_fn.input_blocks.0.0.weight") for k in sd["model"])
    assert set(sd["ema"]) == {n for n, p in r.net.named_parameters() if p.requires_grad}
    r2 = run(train=False)                                              # loads the checkpoint, samples the test set
    out_dir = r2.config.result.sample_to_eval_path
    pngs = [f for _, _, fs in os.walk(out_dir) for f in fs if f.endswith(".png")]
    assert len(pngs) >= 4


def _no_workers(orig):
    def init(self, *a, **kw):
        kw["num_workers"] = 0          # the template asks for 8 loader processes; keep the test light
        return orig(self, *a, **kw)
    return init


def _drop_kw(orig, name):
    def init(self, *a, **kw):
        kw.pop(name, None)
        return orig(self, *a, **kw)
    return init
