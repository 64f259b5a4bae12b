"""Drop-in ``runners.base.EMA`` (namespace-package overlay, like model/BrownianBridge/*.py):
the reference runner's ``from runners.base.EMA import EMA`` (runners/BaseRunner.py:21) resolves here when this repo
is ahead of the BBDM checkout on ``sys.path``.  Same interface as the reference class
(/root/reference/runners/base/EMA.py:4-43); the shadow parameters live in one flat buffer and ``update`` is ONE
multi-tensor kernel launch (``bbdm_ema_multi``, bit-exact with the reference expression) instead of two tensor ops and a
clone per parameter.  This directory deliberately has no ``__init__.py``.
"""
from bbdm_b200.optim import FusedEMA as EMA

__all__ = ["EMA"]
