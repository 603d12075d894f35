"""occlusionenv_b200: B200-native batched implementation of the OcclusionEnv transition
(MILAB-IIT-CV/OcclusionEnv ``environment.py`` step/reset + the vectorised-env interface).

The arithmetic lives in ``libocclb200.so`` (hand-written sm_100a CUDA behind the C-ABI of
``include/occl_b200.h``); this package is the host-side mirror of the reference's interface.
"""
from .config import RasterConfig  # noqa: F401
from .meshes import SceneMesh, default_scene, load_obj, load_teapot, make_box, pack_scene, procedural_scene  # noqa: F401

__all__ = ["RasterConfig", "SceneMesh", "default_scene", "load_obj", "load_teapot", "make_box", "pack_scene",
           "procedural_scene", "OcclusionEnv", "SimpleVecEnv", "BatchedOcclusionVecEnv", "VecEnv"]


def __getattr__(name):
    # torch / CUDA dependent classes are imported lazily so that mesh + config helpers work anywhere
    if name == "OcclusionEnv":
        from .environment import OcclusionEnv
        return OcclusionEnv
    if name in ("SimpleVecEnv", "BatchedOcclusionVecEnv"):
        from . import SubProcVecEnv as m
        return getattr(m, name)
    if name in ("VecEnv", "VecEnvWrapper"):
        from . import baseVecEnv as m
        return getattr(m, name)
    if name == "OcclusionEngine":
        from .engine import OcclusionEngine
        return OcclusionEngine
    raise AttributeError(name)
