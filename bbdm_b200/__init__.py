"""bbdm_b200 -- B200-native (sm_100a) implementation of the BBDM hot path.

Public surface (mirrors the reference):
  model.BrownianBridge.BrownianBridgeModel.BrownianBridgeModel        (drop-in overlay)
  model.BrownianBridge.LatentBrownianBridgeModel.LatentBrownianBridgeModel
  bbdm_b200.unet.UNetModel            -- parameter tree identical to the reference UNetModel
  bbdm_b200.engine.UNetEngine         -- executor over the C-ABI kernels (libbbdm_b200.so)
  bbdm_b200.cabi                      -- ctypes binding of include/bbdm_b200.h
"""
__version__ = "0.1.0"
