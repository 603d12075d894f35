"""Dataset generator outputs (``/root/reference/datasetGenerator.py:68-124``), batched.

The reference renders ``numObj`` runs of ``numFrame`` differentiable steps, one env at a time, and writes per run

    Dataset/run_<i>/RGB/<j>.jpg     img_as_ubyte(obs[..., :3])                                  (:103-105)
    Dataset/run_<i>/Occl/<j>.png    img_as_ubyte(info['full_state'][..., 3])                    (:107-109)
    Dataset/run_<i>/Depth/<j>.png   depth, background -1 -> 0, times 51, truncated to uint8      (:111-114)
    Dataset/run_<i>/params.pickle   flat float64 array of rows [j, elevation, azimuth, g0, g1]   (:99-101, :122-124)

(g = d reward / d action of the step that produced the frame; ``np.append`` flattens the rows, kept).
Here a whole batch of runs advances at once through ``BatchedOcclusionVecEnv``: one differentiable launch chain
per frame for all runs, the files are encoded on the host from one device->host copy per frame.
The encoders (``encode_*``) are pure numpy so that the formats are testable without a GPU.
"""
from __future__ import annotations

import os
import pickle
from typing import Optional

import numpy as np


def img_as_ubyte(x: np.ndarray) -> np.ndarray:
    """skimage.util.img_as_ubyte for a float image: values in [0, 1] -> rint(255 x); negatives clip to 0.
    (skimage raises for |x| > 1; the shaded observation can exceed 1 by an ulp, so this clips instead.)"""
    return np.rint(np.clip(np.asarray(x, np.float64), 0.0, 1.0) * 255.0).astype(np.uint8)


def encode_rgb(obs_chw: np.ndarray) -> np.ndarray:
    """(4,S,S) observation -> (S,S,3) uint8 as the reference hands it to cv2.imwrite (:103-105)."""
    return img_as_ubyte(np.transpose(obs_chw[:3], (1, 2, 0)))


def encode_occlusion(occl_hw: np.ndarray) -> np.ndarray:
    """(S,S) alpha of the occlusion image -> uint8 (:107-109)."""
    return img_as_ubyte(occl_hw)


def encode_depth(obs_chw: np.ndarray) -> np.ndarray:
    """depth channel: background (-1) -> 0, times 51, ``astype('uint8')`` (truncation, wraps above 5.02) (:111-114)."""
    d = np.array(obs_chw[3], np.float32, copy=True)
    d[d == -1] = 0
    d *= 51
    return d.astype(np.int64).astype(np.uint8)  # astype('uint8') of a float: truncate, modulo 256


def write_frame(run_dir: str, j: int, obs_chw: np.ndarray, occl_hw: np.ndarray) -> None:
    import cv2

    cv2.imwrite(os.path.join(run_dir, "RGB", f"{j}.jpg"), encode_rgb(obs_chw))
    cv2.imwrite(os.path.join(run_dir, "Occl", f"{j}.png"), encode_occlusion(occl_hw))
    cv2.imwrite(os.path.join(run_dir, "Depth", f"{j}.png"), encode_depth(obs_chw))


def write_params(run_dir: str, rows: np.ndarray) -> None:
    """rows (numFrame, 5) -> the reference's flat float64 array (np.append without axis flattens)."""
    with open(os.path.join(run_dir, "params.pickle"), "wb") as f:
        pickle.dump(np.asarray(rows, np.float64).reshape(-1), f)


def generate_dataset(out_dir: str = "./Dataset", num_obj: int = 10000, num_frame: int = 20, img_size: int = 512,
                     lr: float = 2.5e-2, azimuth: float = 0.0, batch: int = 256, data=None, seed: Optional[int] = None,
                     device: Optional[str] = None, first_run: int = 0, max_step_factor: int = 5) -> int:
    """Writes runs ``first_run .. first_run + num_obj - 1``; returns the number of frames written.
    Every run: reset(azimuth) then ``num_frame`` differentiable steps with action = lr * randn(2)."""
    import torch

    from .SubProcVecEnv import BatchedOcclusionVecEnv

    gen = torch.Generator().manual_seed(0 if seed is None else int(seed))
    frames = 0
    for start in range(0, num_obj, batch):
        n = min(batch, num_obj - start)
        venv = BatchedOcclusionVecEnv(n, data=data, img_size=img_size, device=device, auto_reset=False)
        venv.reset(azimuth=azimuth)
        dirs = []
        for i in range(n):
            d = os.path.join(out_dir, f"run_{first_run + start + i}")
            for sub in ("Depth", "RGB", "Occl"):
                os.makedirs(os.path.join(d, sub), exist_ok=True)
            dirs.append(d)
        rows = np.zeros((n, num_frame, 5), np.float64)
        eng = venv.engine
        # per-run frame counter: a step whose gradient is not finite is REPEATED with the same action and the frame
        # index does not advance (the reference's `continue`, datasetGenerator.py:93-94), so every run ends with
        # exactly num_frame frames and num_frame rows; runs that are complete keep stepping unobserved
        j_of = np.zeros(n, np.int64)
        action_val = (lr * torch.randn(n, 2, generator=gen)).to(venv.device)
        for _ in range(max_step_factor * num_frame):
            if (j_of >= num_frame).all():
                break
            action = torch.nn.Parameter(action_val.clone())
            obs, reward, finished, info = venv.step(action)
            reward.sum().backward()
            eng.check_status()
            grad = action.grad.detach().cpu().numpy()
            obs_h = obs.detach().cpu().numpy()
            occl_h = eng.occl.detach().cpu().numpy()
            el, az = eng.elevation.cpu().numpy(), eng.azimuth.cpu().numpy()
            ok = np.isfinite(grad).all(axis=1)
            for i in np.nonzero(ok & (j_of < num_frame))[0]:
                j = int(j_of[i])
                rows[i, j] = (j, el[i], az[i], grad[i, 0], grad[i, 1])
                write_frame(dirs[i], j, obs_h[i], occl_h[i])
                frames += 1
            j_of[ok] += 1
            fresh = (lr * torch.randn(n, 2, generator=gen)).to(venv.device)  # a new action only after a kept frame
            action_val = torch.where(torch.from_numpy(ok).to(venv.device)[:, None], fresh, action_val)
        else:
            if not (j_of >= num_frame).all():
                raise RuntimeError("generate_dataset: gradient stayed non-finite for some run")
        for i in range(n):
            write_params(dirs[i], rows[i])
        del venv
    return frames
