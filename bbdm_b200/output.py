"""Output path of ``sample_to_eval``.

The reference converts and writes one image at a time on the sampling thread
(runners/DiffusionBasedModelRunners/BBDMRunner.py:242-253 -> runners/utils.py:67-74: clone, 6 elementwise
launches, a synchronising device->host copy and the PNG encode per image), which serialises with the sampling loop
once the UNet is fast.  Here the whole batch is converted by ONE kernel (``bbdm_denorm_to_uint8``, byte-exact with the
reference expression), copied to pinned host memory asynchronously, and encoded / written by a small pool of
worker threads while the GPU samples the next batch.

    save_single_image(image, save_path, file_name, to_normal=True)     # the reference's signature, same bytes
    save_image_batch(images, save_path, file_names, to_normal=True, writer=None)
    AsyncImageWriter(workers=4).submit(...) / .drain()
    install()      # opt-in: rebind save_single_image in the (unmodified) runner modules to the function above
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor

import torch

from . import cabi

backend_factory = staticmethod(lambda: cabi.CudaBackend())
_be = None


def _backend(t):
    global _be
    if _be is None:
        _be = backend_factory.__func__()
    if not t.is_cuda and getattr(_be, "requires_cuda", True):
        raise RuntimeError("bbdm_b200.output runs only on a CUDA sm_100a device (kernels behind libbbdm_b200.so); "
                           "there is no CPU fallback.")
    return _be


@torch.no_grad()
def images_to_uint8(images: torch.Tensor, to_normal: bool = True) -> torch.Tensor:
    """[B,C,H,W] fp32 -> [B,H,W,C] uint8 on the same device (one launch)."""
    x = images.detach()
    if x.dim() == 3:
        x = x.unsqueeze(0)
    x = x.contiguous().float()
    B, C, H, W = x.shape
    out = torch.empty((B, H, W, C), dtype=torch.uint8, device=x.device)
    _backend(x).denorm_to_uint8(x, bool(to_normal), out)
    return out


def _write_png(arr, path):
    from PIL import Image
    a = arr.numpy()
    Image.fromarray(a[..., 0] if a.shape[-1] == 1 else a).save(path)


class AsyncImageWriter:
    """Pinned staging buffers + worker threads: ``submit`` returns as soon as the device->host copy is enqueued."""

    def __init__(self, workers: int = 4, slots: int = 4):
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.slots = slots
        self._free = threading.Semaphore(slots)
        self._pending = []

    def submit(self, u8: torch.Tensor, paths):
        """u8: [B,H,W,C] uint8 device tensor; paths: B file names."""
        assert u8.dtype == torch.uint8 and u8.dim() == 4 and len(paths) == u8.shape[0]
        self._free.acquire()
        host = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=u8.is_cuda)
        host.copy_(u8, non_blocking=True)
        ev = None
        if u8.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(u8.device))

        def job():
            try:
                if ev is not None:
                    ev.synchronize()
                for i, p in enumerate(paths):
                    _write_png(host[i], p)
            finally:
                self._free.release()
        self._pending.append(self.pool.submit(job))

    def drain(self):
        for f in self._pending:
            f.result()
        self._pending = []

    def close(self):
        self.drain()
        self.pool.shutdown()


@torch.no_grad()
def save_image_batch(images, save_path, file_names, to_normal=True, writer: AsyncImageWriter | None = None):
    u8 = images_to_uint8(images, to_normal)
    paths = [os.path.join(save_path, n) for n in file_names]
    if writer is not None:
        writer.submit(u8, paths)
        return
    host = u8.cpu()
    for i, p in enumerate(paths):
        _write_png(host[i], p)


@torch.no_grad()
def save_single_image(image, save_path, file_name, to_normal=True):
    """Drop-in for runners/utils.py:67-74 (same signature, byte-identical PNG pixels)."""
    save_image_batch(image.unsqueeze(0) if image.dim() == 3 else image, save_path, [file_name], to_normal)


def install():
    """Opt-in: rebind ``save_single_image`` in the reference's runner modules (already imported or not) to the
    kernel-backed function.  The runner source stays untouched."""
    import importlib
    n = 0
    for name in ("runners.utils", "runners.DiffusionBasedModelRunners.BBDMRunner"):
        try:
            mod = importlib.import_module(name)
        except Exception:
            continue
        if hasattr(mod, "save_single_image"):
            mod.save_single_image = save_single_image
            n += 1
    return n
