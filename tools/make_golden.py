"""Generates tests/golden/*.npz from the CPU oracle (oracle/occl_oracle.c), cross-checked against the
independent dense PyTorch formulation (oracle/dense_torch.py) before writing.

The reference itself cannot produce golden vectors here (pytorch3d is not installable offline), so these
pin the oracle against accidental change and give the GPU tests fixed inputs/outputs that travel to the
GPU box.  Run:  python tools/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from occlusionenv_b200.meshes import default_scene  # noqa: E402
from oracle import dense_torch as D  # noqa: E402
from oracle import oracle as O  # noqa: E402

S = 128
POSES = [(0.0, 0.0), (0.7, 0.3), (1.5, 0.0), (float(np.float32(np.pi / 2)), 0.0)]  # (az, el), SURVEY 8c
out_dir = os.path.join(ROOT, "tests", "golden")
os.makedirs(out_dir, exist_ok=True)

for occ in ("teapot", "box"):
    sc = default_scene(occ)
    rec = {}
    for k, (az, el) in enumerate(POSES):
        _, _, C, R, T = O.pose_step(np.zeros(2, np.float32), el, az, 4.0)
        r = O.render_scene(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, S, C, R, T)
        # cross-check with the dense formulation (fp64): alphas within fp32 noise except boundary flips
        verts = torch.tensor(sc.verts, dtype=torch.float64)
        faces = torch.tensor(sc.faces, dtype=torch.long)
        loss_d, alphas_d = D.occlusion_loss(verts, faces, sc.obj_face_start, sc.obj_vert_start, S,
                                            torch.tensor(R, dtype=torch.float64), torch.tensor(T, dtype=torch.float64),
                                            float(O.PROJ_SCALE), float(O.BLUR_RADIUS), float(O.SIGMA), 100)
        err = np.abs(alphas_d.numpy() - r.alphas)
        assert (err > 1e-4).sum() <= 4 and abs(float(loss_d) - float(r.loss)) <= 1e-4 * max(1.0, float(r.loss)), (
            occ, k, err.max(), float(loss_d), float(r.loss))
        rec[f"C{k}"], rec[f"R{k}"], rec[f"T{k}"] = C, R, T
        rec[f"pix_to_face{k}"] = r.pix_to_face.astype(np.int16)
        rec[f"zbuf{k}"] = r.zbuf
        rec[f"alphas{k}"] = r.alphas
        rec[f"rgb{k}"] = r.obs[0]
        rec[f"loss{k}"] = np.float32(r.loss)
        rec[f"nhits_max{k}"] = np.int32(r.nhits.max())
        rec[f"n_covered{k}"] = r.n_covered.astype(np.int32)
        rec[f"n_visible{k}"] = r.n_visible.astype(np.int32)
        print(occ, k, "loss", r.loss, "dense", float(loss_d), "alpha err max", err.max(), "nhits max", r.nhits.max())
    rec["poses"] = np.asarray(POSES, np.float32)
    np.savez_compressed(os.path.join(out_dir, f"scene_{occ}_128.npz"), **rec)

# one short trajectory: reset + 3 steps (state machine, reward, done)
sc = default_scene("teapot")
env = O.OracleOcclusionEnv(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, img_size=64)
env.reset(radius=4.0, azimuth=1.45, elevation=0.1)
acts = np.array([[0.3, -1.0], [1.0, 1.0], [0.0, 0.0], [-0.2, 0.5]], np.float32)
traj = {"actions": acts, "loss0": np.float32(env.fullReward), "mass": np.float32(env.objectMass)}
rews, dones, losses, els, azs = [], [], [], [], []
for a in acts:
    _, r, d, info = env.step(a)
    rews.append(r); dones.append(d); losses.append(info["full_reward"]); els.append(env.elevation); azs.append(env.azimuth)
traj.update(rewards=np.asarray(rews, np.float32), dones=np.asarray(dones), losses=np.asarray(losses, np.float32),
            elevations=np.asarray(els, np.float32), azimuths=np.asarray(azs, np.float32))
np.savez_compressed(os.path.join(out_dir, "trajectory_teapot_64.npz"), **traj)
print("trajectory", traj["rewards"], traj["dones"], traj["losses"])

# camera inside the target's bounding box: clip_faces removes faces and cuts visible ones at z_clip = znear/2
sc = default_scene("teapot")
S2 = 64
C, R, T = O.pose_lookat(1.0, 0.1, 1.0)
r = O.render_scene(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, S2, C, R, T)
assert r.zclip_straddle
np.savez_compressed(os.path.join(out_dir, "scene_teapot_near_64.npz"), C=C, R=R, T=T, pose=np.asarray([1.0, 0.1, 1.0], np.float32),
                    pix_to_face=r.pix_to_face.astype(np.int16), zbuf=r.zbuf, alphas=r.alphas, rgb=r.obs[0],
                    bary=r.bary.astype(np.float32), loss=np.float32(r.loss), nhits=r.nhits.astype(np.int16),
                    n_covered=r.n_covered.astype(np.int32), n_visible=r.n_visible.astype(np.int32))
print("near camera: loss", r.loss, "nhits max", r.nhits.max(), "covered", r.n_covered, "visible", r.n_visible)

# d loss / d (elevation, azimuth) at the poses of the two scene fixtures, float64 autograd of the dense
# formulation; stored next to the fixtures
for occ in ("teapot", "box"):
    sc = default_scene(occ)
    g = []
    for az, el in POSES:
        loss_d, gd = D.loss_and_pose_grad(sc, 128, el, az, 4.0, float(O.PROJ_SCALE), float(O.BLUR_RADIUS), float(O.SIGMA))
        g.append(gd)
        print(occ, "pose", (az, el), "loss", loss_d, "dloss/d(el,az)", gd)
    np.savez_compressed(os.path.join(out_dir, f"grad_{occ}_128.npz"), poses=np.asarray(POSES, np.float32), dloss_del_daz=np.asarray(g))
