"""crimac_unet_b200 — B200-native implementation of the CRIMAC echogram U-Net hot path.

The directory name carries a hyphen (it is the repo's package directory), so it is imported under the alias
``crimac_unet_b200`` (see ``__graft_entry__.load_package()``), or — exactly like the reference's ``crimac_unet/`` —
by putting this directory on ``sys.path`` and doing ``import models.unet as models``
(reference pipeline_train_predict/pipeline.py:32).

Contents: ``csrc/`` hand-written sm_100a kernels + the C-ABI (``include/crimac_b200.h``), ``lib.py``/``engine.py`` the
ctypes host side, ``models/unet.py`` the drop-in nn.Module surface, ``predict.py`` the sliding-window driver,
``train_patches.py`` the on-device training-sample feeder.
"""
__all__ = ["lib", "engine", "models", "predict", "train_patches"]
