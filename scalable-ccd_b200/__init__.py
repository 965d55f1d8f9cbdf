"""B200-native continuous-collision-detection hot path (drop-in for the Scalable-CCD GPU
path).  The product is the CUDA library csrc/ -> libsccd_b200.so behind the C ABI in
include/sccd.h; this package only holds its ctypes binding, the synthetic scene
generators and the multi-GPU orchestration.  Directory name has a hyphen, so import it
through the repo-root helper:  `from _pkg import load_package; sccd = load_package()`.
"""
from . import capi, multigpu, scenes  # noqa: F401
from .capi import Context, SccdError, VF, EE  # noqa: F401

__version__ = "0.1.0"
