"""B200-native rollout hot path for the R2R navigation agents of
IMNearth/Curriculum-Learning-For-VLN (Follower / Self-Monitoring / EnvDrop).

Layout: csrc/ (sm_100a CUDA kernels + the C-ABI of include/vln_b200.h), _lib.py (ctypes
binding), ops.py (autograd wrappers), model/ agent/ engine/ environ/ utils/ (host-side mirror
of the reference's tasks/R2R-judy/src interface for this path).
"""
from . import environ  # noqa: F401
from . import utils  # noqa: F401

__all__ = ["environ", "utils", "ops", "model", "agent", "engine", "compat"]


def __getattr__(name):
    if name in ("ops", "model", "agent", "engine", "compat"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
