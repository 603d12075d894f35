"""Vectorised-environment contract of the reference (``/root/reference/baseVecEnv.py``), re-stated.

Same public names and call semantics so that code written against the reference's ``VecEnv`` keeps
working: ``VecEnv`` (``baseVecEnv.py:57-233``), ``VecEnvWrapper`` (``:236-340``), ``CloudpickleWrapper``
(``:343-356``), ``tile_images`` (``:9-32``) and the two step-protocol errors (``:35-54``).  The reference
never raises the errors; here ``step_wait`` without a pending ``step_async`` does.
"""
from __future__ import annotations

import abc
import inspect
import math
import pickle
from typing import Iterable, List, Optional, Sequence, Union

import numpy as np


def tile_images(img_nhwc):
    """Arrange N images (N,H,W,C) on a near-square grid: rows = ceil(sqrt(N)), cols = ceil(N/rows);
    unused cells are black.  Returns (rows*H, cols*W, C).  (``baseVecEnv.py:9-32``)"""
    imgs = np.asarray(img_nhwc)
    n, h, w, c = imgs.shape
    rows = int(math.ceil(math.sqrt(n)))
    cols = int(math.ceil(n / rows))
    grid = np.zeros((rows * cols, h, w, c), dtype=imgs.dtype)
    grid[:n] = imgs
    return grid.reshape(rows, cols, h, w, c).swapaxes(1, 2).reshape(rows * h, cols * w, c)


class AlreadySteppingError(Exception):
    """``step_async`` called while a step is pending (``baseVecEnv.py:35-43``)."""

    def __init__(self):
        super().__init__("already running an async step")


class NotSteppingError(Exception):
    """``step_wait`` called with no step pending (``baseVecEnv.py:46-54``)."""

    def __init__(self):
        super().__init__("not running an async step")


class VecEnv(abc.ABC):
    """Abstract batch of environments stepped together (``baseVecEnv.py:57-233``)."""

    metadata = {"render.modes": ["human", "rgb_array"]}

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    # -- abstract part of the contract ---------------------------------------------------------
    @abc.abstractmethod
    def reset(self):
        """Reset every environment; returns the stacked observations."""

    @abc.abstractmethod
    def step_async(self, actions):
        """Start a step with one action per environment."""

    @abc.abstractmethod
    def step_wait(self):
        """Finish the pending step: (observations, rewards, dones, infos)."""

    @abc.abstractmethod
    def close(self):
        """Release resources."""

    @abc.abstractmethod
    def get_attr(self, attr_name, indices=None):
        """List of ``attr_name`` over the selected environments."""

    @abc.abstractmethod
    def set_attr(self, attr_name, value, indices=None):
        """Assign ``attr_name`` on the selected environments."""

    @abc.abstractmethod
    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        """Call a method on the selected environments; list of results."""

    @abc.abstractmethod
    def seed(self, seed: Optional[int] = None) -> List[Union[None, int]]:
        """Seed environment i with ``seed + i``; returns what each env's ``seed`` returned."""

    # -- concrete helpers ------------------------------------------------------------------------
    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def get_images(self, *args, **kwargs) -> Sequence[np.ndarray]:
        raise NotImplementedError

    def render(self, mode: str = "human", *args, **kwargs):
        try:
            frames = self.get_images(*args, **kwargs)
        except NotImplementedError:
            print(f"Render not defined for {self}")
            return None
        mosaic = tile_images(frames)
        if mode == "rgb_array":
            return mosaic
        if mode == "human":
            import cv2  # lazy: only needed for on-screen display

            cv2.imshow("vecenv", mosaic[:, :, ::-1])
            cv2.waitKey(1)
            return None
        raise NotImplementedError(mode)

    @property
    def unwrapped(self):
        return self.venv.unwrapped if isinstance(self, VecEnvWrapper) else self

    def getattr_depth_check(self, name, already_found):
        """Name of this class if ``name`` is defined here although an outer wrapper already has it."""
        if already_found and hasattr(self, name):
            return f"{type(self).__module__}.{type(self).__name__}"
        return None

    def _get_indices(self, indices) -> Iterable[int]:
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices


class VecEnvWrapper(VecEnv):
    """Forwards everything to ``venv``; subclasses override what they change (``baseVecEnv.py:236-340``)."""

    def __init__(self, venv, observation_space=None, action_space=None):
        self.venv = venv
        super().__init__(venv.num_envs, observation_space or venv.observation_space,
                         action_space or venv.action_space)
        # every member of the class INCLUDING inherited ones (the reference uses inspect.getmembers, baseVecEnv.py:247)
        self.class_attributes = dict(inspect.getmembers(type(self)))

    def step_async(self, actions):
        self.venv.step_async(actions)

    @abc.abstractmethod
    def reset(self):
        ...

    @abc.abstractmethod
    def step_wait(self):
        ...

    def seed(self, seed=None):
        return self.venv.seed(seed)

    def close(self):
        return self.venv.close()

    def render(self, *args, **kwargs):
        return self.venv.render(*args, **kwargs)

    def get_images(self):
        return self.venv.get_images()

    def get_attr(self, attr_name, indices=None):
        return self.venv.get_attr(attr_name, indices)

    def set_attr(self, attr_name, value, indices=None):
        return self.venv.set_attr(attr_name, value, indices)

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        return self.venv.env_method(method_name, *method_args, indices=indices, **method_kwargs)

    def __getattr__(self, name):
        # only reached when normal lookup fails: search the wrapped chain, refusing shadowed names
        shadow = self.getattr_depth_check(name, already_found=False)
        if shadow is not None:
            me = f"{type(self).__module__}.{type(self).__name__}"
            raise AttributeError(f"Error: Recursive attribute lookup for {name} from {me} is ambiguous "
                                 f"and hides attribute from {shadow}")
        return self.getattr_recursive(name)

    def _get_all_attributes(self):
        attrs = dict(self.__dict__)
        attrs.update(self.class_attributes)
        return attrs

    def getattr_recursive(self, name):
        if name in self._get_all_attributes():
            return getattr(self, name)
        inner = self.__dict__["venv"]
        if hasattr(inner, "getattr_recursive"):
            return inner.getattr_recursive(name)
        return getattr(inner, name)

    def getattr_depth_check(self, name, already_found):
        here = name in self._get_all_attributes()
        if here and already_found:
            return f"{type(self).__module__}.{type(self).__name__}"
        return self.venv.getattr_depth_check(name, already_found or here)


class CloudpickleWrapper:
    """Carries a callable across process boundaries with cloudpickle (``baseVecEnv.py:343-356``)."""

    def __init__(self, var):
        self.var = var

    def __getstate__(self):
        import cloudpickle

        return cloudpickle.dumps(self.var)

    def __setstate__(self, blob):
        self.var = pickle.loads(blob)
