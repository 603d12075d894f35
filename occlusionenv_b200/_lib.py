"""ctypes binding of libocclb200.so (include/occl_b200.h).  No CPU fallback: if the library is missing
or a call fails, this raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

OCCL_ABI_VERSION = 3
OCCL_MAX_OBJ = 4
OCCL_CAM_STRIDE = 48

ST_ZCLIP = 1
ST_KOVERFLOW = 2
ST_HITCAP = 4
ST_OVFCAP = 8
ST_CLIPPED = 16  # informational: faces were cut at z_clip

_ERR = {-1: "OCCL_E_INVALID (bad argument / unsupported configuration)",
        -2: "OCCL_E_CUDA (CUDA call failed)",
        -3: "OCCL_E_SMEM (tile does not fit in shared memory)"}


class OcclConfig(Structure):
    _fields_ = [
        ("image_size", c_int32), ("n_obj", c_int32), ("n_verts", c_int32), ("n_faces", c_int32),
        ("obj_face_start", c_int32 * (OCCL_MAX_OBJ + 1)),
        ("faces_per_pixel", c_int32), ("cull_backfaces", c_int32), ("norm_with_object_size", c_int32),
        ("tile_w", c_int32), ("tile_h", c_int32),
        ("blur_radius", c_float), ("sigma", c_float), ("proj_scale", c_float), ("z_clip", c_float),
        ("step_size", c_float), ("light", c_float * 3),
        ("done_threshold", c_float), ("reward_done", c_float), ("reward_step", c_float),
        ("debug_exact", c_int32), ("ws_budget_mb", c_int32), ("obs_planes", c_int32),
    ]


class OcclScene(Structure):
    _fields_ = [("verts", c_void_p), ("faces", c_void_p), ("verts_env_stride", c_int64),
                ("faces_env_stride", c_int64)]


class OcclState(Structure):
    _fields_ = [("elevation", c_void_p), ("azimuth", c_void_p), ("radius", c_void_p),
                ("full_reward", c_void_p), ("object_mass", c_void_p)]


class OcclWorkspace(Structure):
    _fields_ = [("base", c_void_p), ("bytes", c_size_t)]


class OcclOutputs(Structure):
    _fields_ = [("obs", c_void_p), ("occl", c_void_p), ("reward", c_void_p), ("done", c_void_p),
                ("loss", c_void_p), ("position", c_void_p), ("n_covered", c_void_p),
                ("n_visible", c_void_p), ("status", c_void_p), ("grad_action", c_void_p),
                ("alphas", c_void_p), ("pix_to_face", c_void_p), ("bary", c_void_p), ("nhits", c_void_p),
                ("status_or", c_void_p), ("obs_tile_state", c_void_p)]


OCCL_TILE_STATE_WORDS = 16  # per env: 8 words "tile held a face at the last render" + 8 words of the current skip mask


# every symbol include/occl_b200.h declares
EXPORTS = ["occl_abi_version", "occl_last_cuda_error", "occl_enable_peer_access", "occl_ipc_open", "occl_ipc_close", "occl_ipc_export", "occl_selftest_div", "occl_config_resolve", "occl_workspace_bytes",
           "occl_workspace_offsets",
           "occl_pose_step", "occl_pose_lookat", "occl_pose_set", "occl_project", "occl_raster",
           "occl_finalize", "occl_step", "occl_reset", "occl_render"]

_PKG = os.path.dirname(os.path.abspath(__file__))
# OCCL_B200_LIB: alternative build of the same library (kernel-tuning experiments only)
LIB_PATH = os.environ.get("OCCL_B200_LIB") or os.path.join(_PKG, "libocclb200.so")
_lib = None


class OcclError(RuntimeError):
    pass


def load():
    """Load the in-tree CUDA library; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OcclError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    P = POINTER
    lib.occl_abi_version.restype = c_int
    lib.occl_last_cuda_error.restype = c_char_p
    lib.occl_enable_peer_access.argtypes = [c_int]
    lib.occl_enable_peer_access.restype = c_int
    lib.occl_ipc_open.argtypes = [ctypes.c_char_p, POINTER(c_void_p)]
    lib.occl_ipc_open.restype = c_int
    lib.occl_ipc_export.argtypes = [c_void_p, ctypes.c_char_p, POINTER(c_size_t)]
    lib.occl_ipc_export.restype = c_int
    lib.occl_ipc_close.argtypes = [c_void_p]
    lib.occl_ipc_close.restype = c_int
    lib.occl_selftest_div.argtypes = [ctypes.c_ulonglong, ctypes.c_ulonglong, c_void_p, c_void_p]
    lib.occl_selftest_div.restype = c_int
    lib.occl_config_resolve.argtypes = [P(OcclConfig), c_int]
    lib.occl_config_resolve.restype = c_int
    lib.occl_workspace_bytes.argtypes = [P(OcclConfig), c_int, c_int]
    lib.occl_workspace_bytes.restype = c_size_t
    lib.occl_workspace_offsets.argtypes = [P(OcclConfig), c_int, c_int, P(c_size_t)]
    lib.occl_workspace_offsets.restype = c_int
    lib.occl_pose_step.argtypes = [P(OcclConfig), c_int, c_void_p, OcclState, c_void_p, c_void_p]
    lib.occl_pose_lookat.argtypes = [P(OcclConfig), c_int, OcclState, c_void_p, c_void_p]
    lib.occl_pose_set.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.occl_project.argtypes = [P(OcclConfig), c_int, c_void_p, OcclScene, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.occl_raster.argtypes = [P(OcclConfig), c_int, OcclScene, OcclWorkspace, OcclOutputs, c_void_p]
    lib.occl_finalize.argtypes = [P(OcclConfig), c_int, c_int, c_void_p, OcclState, OcclWorkspace, OcclOutputs, c_void_p]
    lib.occl_step.argtypes = [P(OcclConfig), c_int, c_void_p, OcclScene, OcclState, OcclWorkspace, OcclOutputs, c_void_p]
    lib.occl_reset.argtypes = [P(OcclConfig), c_int, c_void_p, OcclScene, OcclState, OcclWorkspace, OcclOutputs, c_void_p]
    lib.occl_render.argtypes = [P(OcclConfig), c_int, c_void_p, c_void_p, c_void_p, OcclScene, OcclWorkspace,
                                OcclOutputs, c_void_p]
    for name in ("occl_pose_step", "occl_pose_lookat", "occl_pose_set", "occl_project", "occl_raster",
                 "occl_finalize", "occl_step", "occl_reset", "occl_render"):
        getattr(lib, name).restype = c_int
    if lib.occl_abi_version() != OCCL_ABI_VERSION:
        raise OcclError("libocclb200.so ABI version mismatch: rebuild")
    _lib = lib
    return lib


def check(rc: int, where: str):
    if rc != 0:
        detail = ""
        if rc == -2:
            detail = ": " + load().occl_last_cuda_error().decode()
        raise OcclError(f"{where} failed: {_ERR.get(rc, rc)}{detail}")
