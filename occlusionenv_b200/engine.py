"""Device-side engine: owns the HBM-resident scene, per-env state, workspace and output buffers of N
environments and drives the C-ABI (include/occl_b200.h).  PyTorch is used for device memory and
streams only; all arithmetic of the transition runs in libocclb200.so.

Replaces, for N environments at once, what ``OcclusionEnv.reset`` / ``.step`` / ``.render`` do through
pytorch3d (``/root/reference/environment.py:286-396``).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib as L
from .config import RasterConfig
from .meshes import SceneMesh


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class OcclusionEngine:
    def __init__(self, scene, n_envs: int, cfg: RasterConfig, device="cuda:0", debug_outputs: bool = False,
                 per_env_scenes: Optional[list] = None, replicate_scenes: bool = False):
        if not torch.cuda.is_available():
            raise L.OcclError("OcclusionEngine needs a CUDA device (there is no CPU fallback)")
        self.lib = L.load()
        self.device = torch.device(device)
        self.n = int(n_envs)
        self.cfg = cfg
        self.S = cfg.image_size
        rep_idx = None
        if per_env_scenes is not None:
            scene0 = per_env_scenes[0]
            if replicate_scenes:  # env e gets scene e % len(per_env_scenes): replicated on the device, not on the host
                rep_idx = torch.arange(self.n) % len(per_env_scenes)
            else:
                assert len(per_env_scenes) == self.n
            for sc in per_env_scenes:
                assert sc.verts.shape == scene0.verts.shape and sc.faces.shape == scene0.faces.shape
                assert np.array_equal(sc.obj_face_start, scene0.obj_face_start)
            verts = np.stack([sc.verts for sc in per_env_scenes])
            faces = np.stack([sc.faces for sc in per_env_scenes])
            self.scene = scene0
            vstride, fstride = scene0.verts.size, scene0.faces.size
        else:
            self.scene = scene
            verts, faces = scene.verts, scene.faces
            vstride = fstride = 0
        sc = self.scene
        self.n_obj = sc.n_obj
        self.verts = torch.from_numpy(np.ascontiguousarray(verts, np.float32)).to(self.device)
        self.faces = torch.from_numpy(np.ascontiguousarray(faces, np.int32)).to(self.device)
        if rep_idx is not None:
            rep_idx = rep_idx.to(self.device)
            self.verts = self.verts.index_select(0, rep_idx).contiguous()
            self.faces = self.faces.index_select(0, rep_idx).contiguous()
        c = L.OcclConfig()
        c.image_size = cfg.image_size
        c.n_obj = sc.n_obj
        c.n_verts = sc.verts.shape[0]
        c.n_faces = sc.faces.shape[0]
        for i in range(L.OCCL_MAX_OBJ + 1):
            c.obj_face_start[i] = int(sc.obj_face_start[min(i, sc.n_obj)])
        c.faces_per_pixel = cfg.faces_per_pixel
        c.cull_backfaces = int(cfg.cull_backfaces)
        c.norm_with_object_size = int(cfg.norm_with_object_size)
        c.tile_w, c.tile_h = cfg.tile_w, cfg.tile_h
        c.blur_radius = cfg.blur_radius
        c.sigma = cfg.sigma
        c.proj_scale = cfg.proj_scale
        c.z_clip = cfg.z_clip
        c.step_size = cfg.step_size
        for i in range(3):
            c.light[i] = cfg.light[i]
        c.done_threshold = cfg.done_threshold
        c.reward_done = cfg.reward_done
        c.reward_step = cfg.reward_step
        c.debug_exact = int(cfg.debug_exact)
        c.ws_budget_mb = int(cfg.ws_budget_mb)
        c.obs_planes = int(cfg.obs_planes)
        L.check(self.lib.occl_config_resolve(ctypes.byref(c), 1), "occl_config_resolve")
        self.c = c
        self.c_scene = L.OcclScene(self.verts.data_ptr(), self.faces.data_ptr(), vstride, fstride)

        N, S, dev = self.n, self.S, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        # per-env state (environment.py:302-306,323-324)
        self.elevation = torch.zeros(N, **f32)
        self.azimuth = torch.zeros(N, **f32)
        self.radius = torch.full((N,), 4.0, **f32)
        self.full_reward = torch.zeros(N, **f32)
        self.object_mass = torch.ones(N, **f32)
        self.c_state = L.OcclState(self.elevation.data_ptr(), self.azimuth.data_ptr(), self.radius.data_ptr(),
                                   self.full_reward.data_ptr(), self.object_mass.data_ptr())
        # outputs
        self.obs = torch.empty(N, 2 if cfg.obs_planes == 2 else 4, S, S, **f32)
        self.occl = torch.empty(N, S, S, **f32)
        self.reward = torch.zeros(N, **f32)
        self.done = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.loss = torch.zeros(N, **f32)
        self.position = torch.zeros(N, 3, **f32)
        self.n_covered = torch.zeros(N, self.n_obj, dtype=torch.int32, device=dev)
        self.n_visible = torch.zeros(N, self.n_obj, dtype=torch.int32, device=dev)
        self.status = torch.zeros(N, dtype=torch.int32, device=dev)
        self.status_or = torch.zeros(1, dtype=torch.int32, device=dev)  # running OR of all status words (device side)
        self.grad_action = torch.zeros(N, 2, **f32)
        self.debug = debug_outputs
        if debug_outputs:
            self.alphas = torch.empty(N, self.n_obj, S, S, **f32)
            self.pix_to_face = torch.empty(N, S, S, dtype=torch.int32, device=dev)
            self.bary = torch.empty(N, S, S, 3, **f32)
            self.nhits = torch.empty(N, self.n_obj, S, S, dtype=torch.int32, device=dev)
        else:
            self.alphas = self.pix_to_face = self.bary = self.nhits = None
        self._out_cache = {}
        self._tile_state = {}  # destination pointer -> (N, 16) tile state of an incremental delivery (incremental_obs)
        nbytes = self.lib.occl_workspace_bytes(ctypes.byref(c), N, 1)
        if nbytes == 0:
            raise L.OcclError("occl_workspace_bytes returned 0 (invalid configuration)")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.c_ws = L.OcclWorkspace(self.workspace.data_ptr(), nbytes)

    # ------------------------------------------------------------------------------------------
    def outputs(self, with_grad: bool = False, obs: Optional[torch.Tensor] = None, scratch: bool = False) -> L.OcclOutputs:
        """Device pointers of the output buffers.  ``scratch=True`` routes everything except ``obs``
        and ``status`` into throw-away buffers (used by ``render()`` so that the results of the last
        transition stay intact)."""
        key = (with_grad, None if obs is None else obs.data_ptr(), scratch)
        cached = self._out_cache.get(key)
        if cached is not None:
            return cached
        o = L.OcclOutputs()
        self._out_cache[key] = o
        if len(self._out_cache) > 16:
            self._out_cache.pop(next(iter(self._out_cache)))
        o.obs = (self.obs if obs is None else obs).data_ptr()
        o.status_or = self.status_or.data_ptr()
        st = self._tile_state.get(o.obs)
        o.obs_tile_state = st.data_ptr() if st is not None else None
        if scratch:
            if getattr(self, "_scratch", None) is None:
                self._scratch = dict(occl=torch.empty_like(self.occl), loss=torch.empty_like(self.loss),
                                     ncov=torch.empty_like(self.n_covered), nvis=torch.empty_like(self.n_visible))
            sc = self._scratch
            o.occl, o.loss = sc["occl"].data_ptr(), sc["loss"].data_ptr()
            o.n_covered, o.n_visible = sc["ncov"].data_ptr(), sc["nvis"].data_ptr()
            o.status = self.status.data_ptr()
            return o
        o.occl = self.occl.data_ptr()
        o.reward = self.reward.data_ptr()
        o.done = self.done.data_ptr()
        o.loss = self.loss.data_ptr()
        o.position = self.position.data_ptr()
        o.n_covered = self.n_covered.data_ptr()
        o.n_visible = self.n_visible.data_ptr()
        o.status = self.status.data_ptr()
        o.grad_action = self.grad_action.data_ptr() if with_grad else None
        if self.debug:
            o.alphas = self.alphas.data_ptr()
            o.pix_to_face = self.pix_to_face.data_ptr()
            o.bary = self.bary.data_ptr()
            o.nhits = self.nhits.data_ptr()
        return o

    def incremental_obs(self, obs: Optional[torch.Tensor] = None, enable: bool = True) -> Optional[torch.Tensor]:
        """Incremental delivery of the observation into ``obs`` (default: the engine's own buffer; typically a slice of the
        learner's gather buffer in peer memory): a tile that was background at the last render into this destination and
        is background now is not stored again (``OcclOutputs.obs_tile_state``).  The destination stays bit-identical to
        a full write as long as nobody else writes into it -- call again to start over if somebody did.  Returns the
        (N, 16) state tensor that travels with the destination."""
        key = (self.obs if obs is None else obs).data_ptr()
        self._out_cache.clear()
        if not enable:
            self._tile_state.pop(key, None)
            return None
        st = torch.full((self.n, L.OCCL_TILE_STATE_WORDS), -1, dtype=torch.int32, device=self.device)  # 0xFF bytes
        self._tile_state[key] = st
        return st

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_pose(self, radius=None, azimuth=None, elevation=None):
        for dst, src in ((self.radius, radius), (self.azimuth, azimuth), (self.elevation, elevation)):
            if src is not None:
                dst.copy_(torch.as_tensor(src, dtype=torch.float32).to(self.device).expand_as(dst))

    def reset(self, radius=None, azimuth=None, elevation=None, obs: Optional[torch.Tensor] = None,
              mask: Optional[torch.Tensor] = None, scratch_outputs: bool = False):
        """Render half of OcclusionEnv.reset (environment.py:302-328).  ``mask`` (N,) uint8/bool on the
        device restricts the reset (pose values included) to the flagged envs.  ``scratch_outputs=True``
        (auto-reset inside a step): only ``obs`` and the env state are updated, the other outputs of the
        last transition (occlusion map, loss, counts, position, done) stay what the step wrote."""
        if mask is not None:
            assert mask.shape == (self.n,) and mask.is_cuda and mask.is_contiguous()
            mask8 = mask if mask.dtype == torch.uint8 else mask.to(torch.uint8)
            mb = mask8.view(torch.bool)
            for dst, src in ((self.radius, radius), (self.azimuth, azimuth), (self.elevation, elevation)):
                if src is None:
                    continue
                if isinstance(src, (int, float)):
                    dst.masked_fill_(mb, float(src))  # one launch, no host->device copy
                else:
                    v = torch.as_tensor(src, dtype=torch.float32).to(self.device).expand_as(dst)
                    dst.copy_(torch.where(mb, v, dst))
            mask = mask8
        else:
            self.set_pose(radius, azimuth, elevation)
        with torch.cuda.device(self.device):
            L.check(self.lib.occl_reset(ctypes.byref(self.c), self.n, _ptr(mask), self.c_scene, self.c_state, self.c_ws,
                                        self.outputs(False, obs, scratch_outputs), self._stream()), "occl_reset")

    def set_env_scenes(self, env_ids, scenes):
        """Swap the meshes of some envs (per-env scenes only): the ``new_scene`` half of ``OcclusionEnv.reset``
        (``environment.py:292-299``).  Topology sizes (V, F, object ranges) must match the engine's."""
        if self.c_scene.verts_env_stride == 0:
            raise L.OcclError("set_env_scenes needs an engine built with per_env_scenes")
        env_ids = [int(i) for i in env_ids]
        if not env_ids:
            return
        sc0 = self.scene
        for sc in scenes:
            if sc.verts.shape != sc0.verts.shape or sc.faces.shape != sc0.faces.shape or \
                    not np.array_equal(sc.obj_face_start, sc0.obj_face_start):
                raise L.OcclError("per-env scenes must share V, F and the object face ranges")
        idx = torch.as_tensor(env_ids, dtype=torch.long, device=self.device)
        v = torch.from_numpy(np.stack([np.ascontiguousarray(sc.verts, np.float32) for sc in scenes])).to(self.device)
        f = torch.from_numpy(np.stack([np.ascontiguousarray(sc.faces, np.int32) for sc in scenes])).to(self.device)
        self.verts.index_copy_(0, idx, v)
        self.faces.index_copy_(0, idx, f)

    def step(self, action: torch.Tensor, with_grad: bool = False, obs: Optional[torch.Tensor] = None):
        """OcclusionEnv.step (environment.py:352-396) for all envs. action: (N,2) f32 on device."""
        assert action.shape == (self.n, 2) and action.dtype == torch.float32 and action.is_cuda and action.is_contiguous()
        with torch.cuda.device(self.device):
            L.check(self.lib.occl_step(ctypes.byref(self.c), self.n, _ptr(action), self.c_scene, self.c_state, self.c_ws,
                                       self.outputs(with_grad, obs), self._stream()), "occl_step")

    def render(self, R: torch.Tensor, T: torch.Tensor, C: torch.Tensor, obs: Optional[torch.Tensor] = None,
               scratch: bool = False):
        """Render from explicit cameras (environment.py:332-336); no state update."""
        for t, shp in ((R, (self.n, 3, 3)), (T, (self.n, 3)), (C, (self.n, 3))):
            assert tuple(t.shape) == shp and t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
        with torch.cuda.device(self.device):
            L.check(self.lib.occl_render(ctypes.byref(self.c), self.n, _ptr(R), _ptr(T), _ptr(C), self.c_scene, self.c_ws,
                                         self.outputs(False, obs, scratch), self._stream()), "occl_render")

    def step_staged(self, action: torch.Tensor, with_grad: bool = False, raster_events=None):
        """The same transition as ``step`` driven stage by stage through the split C-ABI entry points
        (occl_pose_step -> occl_project -> occl_raster -> occl_finalize).  ``raster_events`` = (start,
        end) CUDA events recorded around the rasteriser launch (bench.py's roofline timing)."""
        lib, c, n = self.lib, ctypes.byref(self.c), self.n
        offs = (ctypes.c_size_t * 4)()
        L.check(lib.occl_workspace_offsets(c, n, int(with_grad), offs), "occl_workspace_offsets")
        base = self.workspace.data_ptr()
        cam, vproj = ctypes.c_void_p(base + offs[0]), ctypes.c_void_p(base + offs[1])
        vtan = ctypes.c_void_p(base + offs[2]) if with_grad else None
        st = self._stream()
        out = self.outputs(with_grad)
        with torch.cuda.device(self.device):
            L.check(lib.occl_pose_step(c, n, _ptr(action), self.c_state, cam, st), "occl_pose_step")
            L.check(lib.occl_project(c, n, cam, self.c_scene, vproj, vtan, _ptr(self.status), st), "occl_project")
            if raster_events is not None:
                raster_events[0].record()
            L.check(lib.occl_raster(c, n, self.c_scene, self.c_ws, out, st), "occl_raster")
            if raster_events is not None:
                raster_events[1].record()
            L.check(lib.occl_finalize(c, n, 0, _ptr(action), self.c_state, self.c_ws, out, st), "occl_finalize")

    def camera_blocks(self) -> torch.Tensor:
        """(N, 48) camera blocks of the last call (R, T, C and their tangents), for tests."""
        n = self.n * L.OCCL_CAM_STRIDE * 4
        return self.workspace[:n].view(torch.float32).view(self.n, L.OCCL_CAM_STRIDE).clone()

    def check_status(self, raise_on=L.ST_ZCLIP | L.ST_HITCAP | L.ST_OVFCAP) -> int:
        """Host-syncing check of the status flags raised since the last check: ONE device word (the kernels keep a
        running OR of the per-env status words, ``OcclOutputs.status_or``), read and cleared here.  Raises on
        conditions the kernels flag instead of computing (selection buffers exceeded).  Per-env detail stays in
        ``self.status`` (flags of the last transition)."""
        val = int(self.status_or.item())  # one 4-byte device->host copy
        if val:
            self.status_or.zero_()
        if val & raise_on:
            raise L.OcclError(f"env status flags set: {val & raise_on:#x} (1 = gradient requested through a face cut at "
                              "z_clip = znear/2 by a build without that path, 4 = more candidate faces on one pixel than "
                              "the top-K selection buffer holds)")
        return val
