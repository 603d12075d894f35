"""Vectorised environments (``/root/reference/SubProcVecEnv.py``).

``SimpleVecEnv`` is the reference's class (``SubProcVecEnv.py:189-285``): a sequential in-process loop
over ``OcclusionEnv`` objects, kept for drop-in use.  ``BatchedOcclusionVecEnv`` is what replaces it on
the hot path: the same ``VecEnv`` contract, but ALL environments advance in one fused launch chain of
libocclb200.so, auto-reset included (``SubProcVecEnv.py:211-214``), with no per-env Python work.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import spaces
from .baseVecEnv import AlreadySteppingError, NotSteppingError, VecEnv
from .config import RasterConfig
from .engine import OcclusionEngine
from .environment import OcclusionEnv, StepFunction, resolve_scene
from .spaces import Box


def copy_obs_dict(obs):
    assert isinstance(obs, OrderedDict), f"unexpected type for observations '{type(obs)}'"
    return OrderedDict((k, v) for k, v in obs.items())


def dict_to_obs(space, obs_dict):
    if isinstance(space, spaces.Dict):
        return obs_dict
    if isinstance(space, spaces.Tuple):
        assert len(obs_dict) == len(space.spaces), "size of observation does not match size of observation space"
        return tuple(obs_dict[i] for i in range(len(space.spaces)))
    assert set(obs_dict.keys()) == {None}, "multiple observation keys for unstructured observation space"
    return obs_dict[None]


def obs_space_info(obs_space):
    """(keys, shapes, dtypes) of a space; unstructured spaces use the single key ``None``
    (``SubProcVecEnv.py:42-69``)."""
    if isinstance(obs_space, spaces.Dict):
        sub = obs_space.spaces
    elif isinstance(obs_space, spaces.Tuple):
        sub = dict(enumerate(obs_space.spaces))
    else:
        assert not hasattr(obs_space, "spaces"), f"Unsupported structured space '{type(obs_space)}'"
        sub = {None: obs_space}
    keys = list(sub.keys())
    return keys, {k: sub[k].shape for k in keys}, {k: sub[k].dtype for k in keys}


class SimpleVecEnv(VecEnv):
    """Sequential loop over independent ``OcclusionEnv`` objects -- the reference's semantics exactly,
    including its quirks: ``reset`` draws azimuth from U(-40, 40) radians and
    returns (N,1,4,S,S); a finished env is reset with defaults and its last observation is stored in
    ``info['terminal_observation']``."""

    def __init__(self, env_fns):
        self.envs = [fn() for fn in env_fns]
        env = self.envs[0]
        super().__init__(len(env_fns), env.observation_space, env.action_space)
        self.keys, _, _ = obs_space_info(env.observation_space)
        self.actions = None

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        if self.actions is None:
            raise NotSteppingError()
        obs_buf, rews, dones, infos = [], [], [], []
        for i, env in enumerate(self.envs):
            obs, rew, done, info = env.step(self.actions[i])
            if done:
                info["terminal_observation"] = obs
                obs = env.reset()
            obs_buf.append(obs.squeeze())
            rews.append(rew)
            dones.append(done)
            infos.append(info)
        self.actions = None
        return torch.stack(obs_buf), torch.stack(rews), torch.stack(dones), infos

    def seed(self, seed=None):
        return [env.seed(seed + i) for i, env in enumerate(self.envs)]

    def reset(self):
        rng = np.random.default_rng()
        return torch.stack([env.reset(azimuth=rng.uniform(low=-40, high=40)) for env in self.envs])

    def close(self):
        for env in self.envs:
            env.close()

    def get_images(self) -> Sequence[np.ndarray]:
        out = []
        for env in self.envs:
            rgba, _ = env.render()
            out.append(rgba[0, ..., :3].detach().cpu().numpy())
        return out

    def render(self, mode: str = "human"):
        if self.num_envs == 1 and mode == "human":
            prev, self.envs[0].renderMode = self.envs[0].renderMode, "human"
            try:
                return self.envs[0].render()
            finally:
                self.envs[0].renderMode = prev
        return super().render(mode=mode)

    def get_attr(self, attr_name, indices=None):
        return [getattr(e, attr_name) for e in self._get_target_envs(indices)]

    def set_attr(self, attr_name, value, indices=None):
        for e in self._get_target_envs(indices):
            setattr(e, attr_name, value)

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        return [getattr(e, method_name)(*method_args, **method_kwargs) for e in self._get_target_envs(indices)]

    def _get_target_envs(self, indices):
        return [self.envs[i] for i in self._get_indices(indices)]


class LazyInfos(Sequence):
    """``infos`` of a batched step: behaves like the reference's ``list[dict]`` but builds a dict only
    when an element is read (65 536 Python dicts per step would dominate the step time).  Everything in it is
    the TERMINAL step's result, also for envs that the same call auto-reset (``SubProcVecEnv.py:210-214``): the
    masked reset renders into scratch outputs, so the engine's occlusion map / loss / counts / position are
    still the step's.  ``full_state`` is a view of the engine's occlusion map, valid until the next step."""

    def __init__(self, venv, terminal_obs=None, done=None):
        self._v = venv
        self._occl = venv.engine.occl
        self._pos = venv.engine.position.clone()
        self._loss = venv.engine.loss.clone()
        self._ncov = venv.engine.n_covered.clone()
        self._nvis = venv.engine.n_visible.clone()
        self._term = terminal_obs
        self._done = done

    def __len__(self):
        return self._v.num_envs

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        occl = self._occl[i]
        n_pairs = self._v.engine.n_obj * (self._v.engine.n_obj - 1) // 2
        full = torch.cat([torch.full(occl.shape + (3,), float(n_pairs), device=occl.device), occl[..., None]], -1)
        d = {"full_state": full[None], "position": self._pos[i], "full_reward": self._loss[i],
             "n_covered": self._ncov[i], "n_visible": self._nvis[i]}
        if self._term is not None and bool(self._done[i]):
            d["terminal_observation"] = self._term[i][None]
        return d


class BatchedOcclusionVecEnv(VecEnv):
    """N occlusion environments advanced by one fused GPU launch chain.

    * ``step(actions)``: actions (N,2) tensor (host or device; may require grad) ->
      ``(obs (N,4,S,S), rewards (N,), dones (N,) bool, infos)`` like ``SimpleVecEnv.step_wait``
      (``SubProcVecEnv.py:203-220``).  Finished envs are reset on the device in the same call
      (defaults radius 4, azimuth 0, elevation 0, as ``SubProcVecEnv.py:214``); set
      ``keep_terminal_obs=True`` to get ``info['terminal_observation']`` (costs one host sync).
    * ``reset()``: azimuth ~ U(-40, 40) radians per env (the reference's range, ``SubProcVecEnv.py:233``),
      from a seedable generator; returns (N,4,S,S).
    * ``scene_sampler``: a callable returning a ``SceneMesh`` (stand-in for the ShapeNet loader,
      ``environment.py:91-198``) gives every env its OWN meshes and the reference's ``new_scene`` behaviour
      (``environment.py:292-298,327-328``): ``reset()`` draws a new scene per env and redraws, up to 10 times,
      the envs whose reset render shows no occlusion (loss <= 0.1); with ``resample_on_auto_reset=True`` the
      auto-reset inside ``step`` does the same for the finished envs (one host sync per step to learn which).
    * Returned tensors are views of the engine's output buffers, valid until the next call
      (pass ``copy_outputs=True`` to get fresh tensors like the reference).
    """

    def __init__(self, num_envs: int, data=None, img_size: int = 512, device: Optional[str] = None,
                 auto_reset: bool = True, keep_terminal_obs: bool = False, copy_outputs: bool = False,
                 reset_azimuth_range=(-40.0, 40.0), cfg: Optional[RasterConfig] = None, env_offset: int = 0,
                 per_env_scenes: Optional[list] = None, scene_sampler=None, resample_on_auto_reset: bool = False,
                 max_resets: int = 10):
        self.img_size = img_size
        self.device = torch.device(device or "cuda:0")
        cfg = cfg or RasterConfig(image_size=img_size)
        if cfg.image_size != img_size:
            raise ValueError("cfg.image_size and img_size disagree")
        self.scene_sampler = scene_sampler
        self.resample_on_auto_reset = bool(resample_on_auto_reset) and scene_sampler is not None
        self.max_resets = int(max_resets)
        if scene_sampler is not None and per_env_scenes is None:
            per_env_scenes = [scene_sampler() for _ in range(num_envs)]
        scene = per_env_scenes[0] if per_env_scenes is not None else resolve_scene(data)
        self.engine = OcclusionEngine(scene, num_envs, cfg, device=str(self.device), per_env_scenes=per_env_scenes)
        super().__init__(num_envs, Box(0, 1, shape=(4, img_size, img_size)), Box(low=-0.1, high=0.1, shape=(2,)))
        self.auto_reset = auto_reset
        self.keep_terminal_obs = keep_terminal_obs
        self.copy_outputs = copy_outputs
        self.reset_azimuth_range = reset_azimuth_range
        self.env_offset = env_offset  # global id of env 0 (multi-GPU sharding)
        self.actions = None
        self._gen = torch.Generator().manual_seed(0)
        self.step_size = cfg.step_size
        self.normWithObjectSize = cfg.norm_with_object_size
        self.renderMode = ""

    # -- VecEnv contract ---------------------------------------------------------------------------
    def seed(self, seed: Optional[int] = None):
        if seed is None:
            seed = int(np.random.default_rng().integers(2 ** 31))
        self._gen.manual_seed(int(seed) + self.env_offset)
        return [seed + self.env_offset + i for i in range(self.num_envs)]

    def reset(self, radius=4.0, azimuth=None, elevation=0.0, new_scene: bool = True):
        if azimuth is None:
            lo, hi = self.reset_azimuth_range
            azimuth = lo + (hi - lo) * torch.rand(self.num_envs, generator=self._gen)
        eng = self.engine
        if self.scene_sampler is not None and new_scene:
            eng.set_env_scenes(range(self.num_envs), [self.scene_sampler() for _ in range(self.num_envs)])
        eng.reset(radius=radius, azimuth=azimuth, elevation=elevation)
        if self.scene_sampler is not None and new_scene:
            self._redraw_unoccluded(torch.ones(self.num_envs, dtype=torch.bool, device=self.device))
        self.actions = None
        return self._out(eng.obs)

    def _redraw_unoccluded(self, candidates: torch.Tensor, scratch_outputs: bool = False):
        """``environment.py:327-328``: an env whose reset render shows no occlusion (loss <= 0.1) gets another
        scene, at most ``max_resets`` renders in all; the pose stays what the reset set."""
        eng = self.engine
        for _ in range(self.max_resets - 1):
            again = candidates & ~(eng.full_reward > eng.c.done_threshold)
            ids = torch.nonzero(again).flatten().tolist()  # host sync, as the reference's `if loss > 0.1`
            if not ids:
                return
            eng.set_env_scenes(ids, [self.scene_sampler() for _ in ids])
            eng.reset(mask=again.to(torch.uint8), scratch_outputs=scratch_outputs)
            candidates = again

    def step_async(self, actions):
        if self.actions is not None:
            raise AlreadySteppingError()
        self.actions = actions

    def step_wait(self):
        if self.actions is None:
            raise NotSteppingError()
        actions, self.actions = self.actions, None
        eng = self.engine
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions), dtype=torch.float32)
        if actions.requires_grad and torch.is_grad_enabled():
            rewards = StepFunction.apply(actions, eng)
        else:
            a = actions.detach().to(device=self.device, dtype=torch.float32, non_blocking=True).reshape(self.num_envs, 2)
            eng.step(a.contiguous())
            rewards = self._out(eng.reward)
        dones = eng.done.bool()
        if self.auto_reset:
            infos = LazyInfos(self, eng.obs.clone() if self.keep_terminal_obs else None, dones)
            # masked device-side reset of the finished envs; `done` itself is the mask.  Only obs and the env state
            # change: the step's occlusion map, loss, counts, position and done stay in place for `infos`.
            if self.resample_on_auto_reset:
                ids = torch.nonzero(dones).flatten().tolist()
                if ids:
                    eng.set_env_scenes(ids, [self.scene_sampler() for _ in ids])
            eng.reset(radius=4.0, azimuth=0.0, elevation=0.0, mask=eng.done, scratch_outputs=True)
            if self.resample_on_auto_reset and ids:
                self._redraw_unoccluded(dones, scratch_outputs=True)
        else:
            infos = LazyInfos(self, None, None)
        return self._out(eng.obs), rewards, dones, infos

    def check_status(self, raise_on=L.ST_ZCLIP | L.ST_HITCAP | L.ST_OVFCAP) -> int:
        """Flags raised by ANY env in ANY transition since the last check (the kernels keep a running OR on the
        device, so the hot loop never syncs for it); raises ``OcclError`` on the ones that mean a wrong result."""
        return self.engine.check_status(raise_on)

    def close(self):
        pass

    def _out(self, t):
        return t.clone() if self.copy_outputs else t

    # -- attribute plumbing: the batch is homogeneous, so attributes live on this object ------------
    def get_attr(self, attr_name, indices=None):
        return [getattr(self, attr_name) for _ in self._get_indices(indices)]

    def set_attr(self, attr_name, value, indices=None):
        setattr(self, attr_name, value)
        if attr_name == "step_size":
            self.engine.c.step_size = float(value)
        if attr_name == "normWithObjectSize":
            self.engine.c.norm_with_object_size = int(bool(value))

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        if method_name in ("close", "detach"):
            return [None for _ in self._get_indices(indices)]
        raise NotImplementedError(f"env_method({method_name!r}) has no per-env meaning on the batched env")

    def get_images(self) -> Sequence[np.ndarray]:
        rgb = self.engine.obs[:, :3].permute(0, 2, 3, 1)
        return list(rgb.detach().cpu().numpy())
