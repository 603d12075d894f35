"""Drop-in ``OcclusionEnv`` (``/root/reference/environment.py:201-402``) on the B200 engine.

Same constructor, attributes and ``reset`` / ``step`` / ``render`` / ``seed`` / ``close`` / ``detach``
contract; the arithmetic (pose -> projection -> rasterisation -> soft silhouettes -> flat-shaded RGBD ->
occlusion reward [-> gradient to the action]) runs in libocclb200.so through ``OcclusionEngine``.
A single environment is the N=1 view of the batched engine; use
``SubProcVecEnv.BatchedOcclusionVecEnv`` for throughput.

Deliberate differences: the reward is the n-object form
``sum_{i<j} A_i A_j`` (B-1: the reference indexes a 4th mesh its default loader never returns);
no bare ``except`` retry (a failure raises); ``data`` may be ``None`` (teapot + teapot shifted by +2 in x,
``environment.py:53-88``), ``"box"`` (teapot + box occluder), a ``SceneMesh``, or a callable returning a
``SceneMesh`` (a scene sampler standing in for the ShapeNet loader, ``environment.py:91-198``).
"""
from __future__ import annotations

import random
from typing import Callable, Optional, Union

import numpy as np
import torch

from . import _lib as L
from .config import RasterConfig
from .engine import OcclusionEngine
from .meshes import SceneMesh, default_scene
from .spaces import Box


def look_at_rotation_torch(camera_position: torch.Tensor):
    """pytorch3d ``look_at_rotation`` (at = origin, up = +y) and ``T = -R^T C`` for (N,3) positions;
    used only by ``render()`` (``environment.py:334-335``), which takes an explicit position."""
    C = camera_position.to(torch.float32)

    def nrm(v, eps=1e-5):
        return v / v.norm(dim=1, keepdim=True).clamp_min(eps)

    z = nrm(-C)
    up = torch.tensor([[0.0, 1.0, 0.0]], device=C.device).expand_as(C)
    x = nrm(torch.cross(up, z, dim=1))
    y = nrm(torch.cross(z, x, dim=1))
    close = (x.abs() <= 5e-3).all(dim=1, keepdim=True)
    x = torch.where(close, nrm(torch.cross(y, z, dim=1)), x)
    R = torch.stack([x, y, z], dim=2)  # axes as columns
    T = -torch.bmm(R.transpose(1, 2), C[:, :, None])[:, :, 0]
    return R.contiguous(), T.contiguous()


class StepFunction(torch.autograd.Function):
    """Differentiable transition: forward runs the fused step with forward-mode tangents, so the
    gradient d reward / d action is available immediately; backward scales it by the incoming
    gradient (``demo.py:85-86``, ``train_predict.py:52``)."""

    @staticmethod
    def forward(ctx, action: torch.Tensor, engine: OcclusionEngine):
        a = action.detach().to(device=engine.device, dtype=torch.float32).reshape(engine.n, 2).contiguous()
        engine.step(a, with_grad=True)
        ctx.save_for_backward(engine.grad_action.clone())
        ctx.action_shape = action.shape
        ctx.action_device = action.device
        ctx.action_dtype = action.dtype
        return engine.reward.clone()

    @staticmethod
    def backward(ctx, grad_reward):
        (ga,) = ctx.saved_tensors
        g = grad_reward.reshape(-1, 1).to(ga.dtype) * ga
        return g.reshape(ctx.action_shape).to(device=ctx.action_device, dtype=ctx.action_dtype), None


def resolve_scene(data) -> SceneMesh:
    if data is None:
        return default_scene("teapot")
    if isinstance(data, str):
        return default_scene(data)
    if isinstance(data, SceneMesh):
        return data
    if callable(data):
        sc = data()
        if not isinstance(sc, SceneMesh):
            raise TypeError("scene sampler must return a SceneMesh")
        return sc
    raise TypeError(f"unsupported scene source {type(data)!r}")


class OcclusionEnv:
    def __init__(self, data: Union[None, str, SceneMesh, Callable[[], SceneMesh]] = None, img_size: int = 512,
                 device: Optional[str] = None):
        self.metadata = {"render.modes": ["human", "rgb_array"]}
        self.normWithObjectSize = False
        self.img_size = img_size
        if not torch.cuda.is_available():
            raise L.OcclError("OcclusionEnv needs a CUDA device: the transition runs in libocclb200.so "
                              "(no CPU fallback)")
        self.device = torch.device(device or "cuda:0")
        self.shapenet_dataset = data  # reference attribute name for the scene source
        self.step_size = 0.05
        self.observation_space = Box(0, 1, shape=(4, img_size, img_size))
        self.action_space = Box(low=-0.1, high=0.1, shape=(2,))
        self.renderMode = ""  # 'human'
        # status flags (selection buffers exceeded, ...) are ORed on the device; reading them costs one 4-byte
        # device->host copy per call.  Set to False in a tight loop and call ``check_status()`` when convenient.
        self.check_status_every_step = True
        self.image = None
        self.meshes = None
        self._engine: Optional[OcclusionEngine] = None
        self._scene: Optional[SceneMesh] = None
        self.camera_position = torch.zeros(3, device=self.device)

    # -- reference attributes backed by engine state -------------------------------------------
    @property
    def elevation(self):
        return self._engine.elevation

    @property
    def azimuth(self):
        return self._engine.azimuth

    @property
    def radius(self):
        return self._engine.radius

    @property
    def fullReward(self):
        return self._engine.full_reward[0]

    @property
    def objectMass(self):
        return self._engine.object_mass[0]

    def seed(self, seed):
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        torch.cuda.manual_seed(seed)

    # -------------------------------------------------------------------------------------------
    def _ensure_engine(self, scene: SceneMesh):
        same = (self._engine is not None and self._scene is not None
                and self._scene.verts.shape == scene.verts.shape and self._scene.faces.shape == scene.faces.shape
                and np.array_equal(self._scene.obj_face_start, scene.obj_face_start)
                and self._engine.S == self.img_size)
        if same:
            if scene is not self._scene:
                self._engine.verts.copy_(torch.from_numpy(scene.verts))
                self._engine.faces.copy_(torch.from_numpy(scene.faces))
        else:
            cfg = RasterConfig(image_size=self.img_size, step_size=self.step_size,
                               norm_with_object_size=self.normWithObjectSize)
            self._engine = OcclusionEngine(scene, 1, cfg, device=str(self.device))
        self._scene = scene
        dev = self.device
        self.meshes = [(torch.from_numpy(scene.verts).to(dev), torch.from_numpy(scene.faces).to(dev))]
        for i in range(scene.n_obj):
            v, f = scene.object(i)
            self.meshes.append((torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)))

    def _sync_knobs(self):
        self._engine.c.step_size = float(self.step_size)
        self._engine.c.norm_with_object_size = int(bool(self.normWithObjectSize))

    def _full_state(self) -> torch.Tensor:
        """``self.image`` (``environment.py:373``): RGB channels are the constant sums of products of the
        silhouettes' all-ones RGB; the alpha channel carries the occlusion map."""
        occl = self._engine.occl.clone()
        n_pairs = self._engine.n_obj * (self._engine.n_obj - 1) // 2
        rgb = torch.full(occl.shape + (3,), float(n_pairs), device=occl.device)
        return torch.cat([rgb, occl[..., None]], dim=-1)

    def reset(self, new_scene=True, radius=4.0, azimuth=0.0, elevation=0.0):
        max_resets = 10
        resets = 0
        while True:
            resets += 1
            if new_scene or self._engine is None:
                scene = resolve_scene(self.shapenet_dataset)
                if scene.max_object_faces > 250000:  # environment.py:296-298
                    if callable(self.shapenet_dataset) and resets < max_resets:
                        continue
                    raise L.OcclError("mesh too large (> 250000 faces per object)")
                self._ensure_engine(scene)
            self._sync_knobs()
            self.camera_position = torch.zeros(3, device=self.device)  # environment.py:302
            self._engine.reset(radius=float(radius), azimuth=float(azimuth), elevation=float(elevation))
            self._engine.check_status()
            self.image = self._full_state()
            occluded = bool(self._engine.loss[0] > 0.1)
            resample = callable(self.shapenet_dataset) and new_scene
            if occluded or resets == max_resets or not resample:
                return self._engine.obs.clone()

    def step(self, action):
        if self._engine is None:
            raise L.OcclError("step() called before reset()")
        self.detach()
        self._sync_knobs()
        action = torch.as_tensor(action, dtype=torch.float32) if not torch.is_tensor(action) else action
        eng = self._engine
        if action.requires_grad and torch.is_grad_enabled():
            reward = StepFunction.apply(action, eng)[0]
        else:
            eng.step(action.detach().to(device=self.device, dtype=torch.float32).reshape(1, 2).contiguous())
            reward = eng.reward.clone()[0]
        if self.check_status_every_step:
            eng.check_status()
        observation = eng.obs.clone()
        self.camera_position = eng.position[0].clone()
        self.image = self._full_state()
        finished = eng.done.bool()[0].clone()
        info = {"full_state": self.image, "position": self.camera_position, "full_reward": eng.loss[0].clone(),
                "n_covered": eng.n_covered[0].clone(), "n_visible": eng.n_visible[0].clone()}
        return observation, reward, finished, info

    def render(self):
        """``environment.py:332-347``: re-render the observation from ``self.camera_position``."""
        C = self.camera_position[None, :].to(self.device).contiguous()
        R, T = look_at_rotation_torch(C)
        obs = torch.empty_like(self._engine.obs)
        self._engine.render(R, T, C, obs=obs, scratch=True)
        depth = obs[:, 3:4].permute(0, 2, 3, 1).contiguous()
        rgba = torch.cat([obs[:, :3], (obs[:, 3:4] >= 0).to(obs.dtype)], dim=1).permute(0, 2, 3, 1).contiguous()
        if self.renderMode == "human":
            import cv2

            img = rgba.detach().squeeze().cpu().numpy()[..., :3]
            d = depth.detach().squeeze().cpu().numpy().copy()
            d[d == -1] = 0
            cv2.imshow("Environment", img)
            cv2.imshow("Environment Depth", (d * 51).astype("uint8"))
            cv2.waitKey(25)
            return None
        return rgba, depth

    def check_status(self) -> int:
        """Flags raised since the last check (one device word); raises ``OcclError`` on the ones that mean a wrong result."""
        return self._engine.check_status()

    def close(self):
        pass

    def detach(self):
        # engine state never carries autograd history; kept for API compatibility (environment.py:398-402)
        self.camera_position = self.camera_position.detach()
