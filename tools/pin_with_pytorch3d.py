#!/usr/bin/env python
"""Pin the oracle against the REAL pytorch3d -- to be run on any machine that has pytorch3d 0.6.2 / 0.7.x installed.

This container (and the GPU box) has no pytorch3d and no network, so parity is "unpinned": the CUDA kernels are
compared with oracle/, and oracle/ restates pytorch3d from its published algorithm.  This script closes the loop in
one command wherever the library exists:

    python tools/pin_with_pytorch3d.py [--size 128] [--device cpu]

It builds the reference's renderers exactly as ``createRenderers`` does (``/root/reference/environment.py:234-284``:
FoVPerspectiveCameras defaults, BlendParams(sigma=1e-4, gamma=1e-4), RasterizationSettings(blur_radius = ln(9999) *
1e-4, faces_per_pixel=100 / 1, cull_backfaces=True), SoftSilhouetteShader, HardFlatShader + PointLights((2,2,-2))),
renders the golden poses of tests/golden/scene_{teapot,box}_128.npz (+ the near-camera clip fixture) through
pytorch3d's MeshRasterizer on the CPU (naive rasteriser, the path BASELINE config 1 names), and compares

    pix_to_face (K=1)         bit-exact        zbuf (K=1)            bit-exact
    per-object alpha          1e-5 relative    flat-shaded RGB       1e-5 relative
    soft hits per pixel       exact            loss                  1e-5 relative

first with the committed golden fixtures, then with the oracle re-run under each of its DISCRETIONARY switches
(oracle/oracle.py::set_option; DESIGN.md section 3 lists them): whichever switch removes a mismatch names the choice
that differs from the library.  Exit code 0 = everything within tolerance with the default switches ("pinned").
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def p3d_render(scene, S, R, T, device="cpu"):
    """The reference's two renderers on one pose; returns dict of numpy arrays."""
    import torch
    from pytorch3d.renderer import (BlendParams, FoVPerspectiveCameras, HardFlatShader, MeshRasterizer, MeshRenderer,
                                    PointLights, RasterizationSettings, SoftSilhouetteShader, TexturesVertex)
    from pytorch3d.structures import Meshes

    dev = torch.device(device)
    cameras = FoVPerspectiveCameras(device=dev)                                   # environment.py:238
    blend = BlendParams(sigma=1e-4, gamma=1e-4)                                   # :242
    n_faces = int(scene.max_object_faces)
    rs_soft = RasterizationSettings(image_size=S, blur_radius=np.log(1.0 / 1e-4 - 1.0) * blend.sigma, faces_per_pixel=100,
                                    cull_backfaces=True, max_faces_per_bin=max(n_faces, 10000))   # :249-255
    rs_hard = RasterizationSettings(image_size=S, blur_radius=0.0, faces_per_pixel=1, cull_backfaces=True,
                                    max_faces_per_bin=max(n_faces, 10000))                        # :267-273
    soft_raster = MeshRasterizer(cameras=cameras, raster_settings=rs_soft)
    sil = MeshRenderer(rasterizer=soft_raster, shader=SoftSilhouetteShader(blend_params=blend))  # :258-264
    lights = PointLights(device=dev, location=((2.0, 2.0, -2.0),))                               # :275
    hard_raster = MeshRasterizer(cameras=cameras, raster_settings=rs_hard)
    flat = HardFlatShader(device=dev, cameras=cameras, lights=lights)                            # :283

    def mesh(v, f):
        v = torch.tensor(v, dtype=torch.float32, device=dev)
        f = torch.tensor(f, dtype=torch.int64, device=dev)
        return Meshes(verts=[v], faces=[f], textures=TexturesVertex(verts_features=torch.ones_like(v)[None]))

    Rt = torch.tensor(R[None], dtype=torch.float32, device=dev)
    Tt = torch.tensor(T[None], dtype=torch.float32, device=dev)
    out = {"alphas": [], "nhits": []}
    for i in range(scene.n_obj):
        m = mesh(*scene.object(i))
        frags = soft_raster(m, R=Rt, T=Tt)
        out["nhits"].append((frags.pix_to_face[0] >= 0).sum(-1).cpu().numpy())      # after the K cut: min(hits, K)
        out["alphas"].append(sil(meshes_world=m, R=Rt, T=Tt)[0, ..., 3].cpu().numpy())
    full = mesh(scene.verts, scene.faces)
    frags = hard_raster(full, R=Rt, T=Tt)
    img = flat(frags, full, cameras=cameras, lights=lights)      # (1,S,S,4); cameras carry R, T from the rasteriser call
    out["pix_to_face"] = frags.pix_to_face[0, ..., 0].cpu().numpy().astype(np.int64)
    out["zbuf"] = frags.zbuf[0, ..., 0].cpu().numpy()
    out["rgb"] = img[0, ..., 0].cpu().numpy()
    out["alphas"] = np.stack(out["alphas"])
    out["nhits"] = np.stack(out["nhits"])
    occl = np.zeros_like(out["alphas"][0])
    for i in range(scene.n_obj):
        for j in range(i + 1, scene.n_obj):
            occl = occl + out["alphas"][i] * out["alphas"][j]
    out["loss"] = float(np.sum(occl.astype(np.float64) ** 2))
    return out


def compare(tag, got, want, K=100):
    """`got`: oracle / fixture values, `want`: pytorch3d.  Returns the number of failed criteria."""
    bad = 0

    def line(name, ok, detail):
        nonlocal bad
        bad += 0 if ok else 1
        print(f"  [{'ok' if ok else 'MISMATCH'}] {tag:28s} {name:14s} {detail}")

    d = got["pix_to_face"].astype(np.int64) != want["pix_to_face"]
    line("pix_to_face", not d.any(), f"{int(d.sum())} px differ")
    d = got["zbuf"] != want["zbuf"]
    line("zbuf", not d.any(), f"{int(d.sum())} px differ, max |d| {np.abs(got['zbuf'] - want['zbuf']).max():.3e}")
    a = np.abs(got["alphas"] - want["alphas"])
    tol = 1e-5 * np.abs(want["alphas"]) + 2e-6
    line("alpha", bool((a <= tol).all()), f"max |d| {a.max():.3e}, {int((a > tol).sum())} px over 1e-5 rel + 2e-6")
    if "nhits" in got:
        nh = np.minimum(got["nhits"], K) != want["nhits"]
        line("soft hits", not nh.any(), f"{int(nh.sum())} px differ")
    r = np.abs(got["rgb"] - want["rgb"])
    line("rgb", bool((r <= 1e-5 * np.abs(want["rgb"]) + 1e-6).all()), f"max |d| {r.max():.3e}")
    rel = abs(got["loss"] - want["loss"]) / max(abs(want["loss"]), 1e-6)
    line("loss", rel <= 1e-5 or abs(got["loss"] - want["loss"]) <= 1e-6, f"{got['loss']:.6f} vs {want['loss']:.6f}")
    return bad


def oracle_render(O, scene, S, C, R, T):
    o = O.render_scene(scene.verts, scene.faces, scene.obj_face_start, scene.obj_vert_start, S, C, R, T)
    return {"pix_to_face": o.pix_to_face, "zbuf": o.zbuf, "alphas": o.alphas, "nhits": o.nhits, "rgb": o.obs[0],
            "loss": float(o.loss)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    args = ap.parse_args()
    try:
        import pytorch3d  # noqa: F401
    except ImportError:
        print("pytorch3d is not installed here: nothing can be pinned (parity stays 'unpinned', DESIGN.md section 3)")
        return 2
    from occlusionenv_b200.meshes import default_scene
    from oracle import oracle as O

    cases = []
    gold = os.path.join(ROOT, "tests", "golden")
    for occ in ("teapot", "box"):
        g = np.load(os.path.join(gold, f"scene_{occ}_128.npz"))
        for k in range(len(g["poses"])):
            fix = {"pix_to_face": g[f"pix_to_face{k}"], "zbuf": g[f"zbuf{k}"], "alphas": g[f"alphas{k}"],
                   "rgb": g[f"rgb{k}"], "loss": float(g[f"loss{k}"])}
            cases.append((f"{occ} pose {k} 128^2", default_scene(occ), 128, g[f"C{k}"], g[f"R{k}"], g[f"T{k}"], fix))
    g = np.load(os.path.join(gold, "scene_teapot_near_64.npz"))
    fix = {"pix_to_face": g["pix_to_face"], "zbuf": g["zbuf"], "alphas": g["alphas"], "nhits": g["nhits"], "rgb": g["rgb"],
           "loss": float(g["loss"])}
    cases.append(("teapot near camera (clip) 64^2", default_scene("teapot"), 64, g["C"], g["R"], g["T"], fix))

    failed_default = 0
    for tag, scene, S, C, R, T, fix in cases:
        print(tag)
        want = p3d_render(scene, S, R, T, args.device)
        failed_default += compare("committed golden fixture", fix, want)
        base = compare("oracle, default switches", oracle_render(O, scene, S, C, R, T), want)
        failed_default += base
        if base:
            for name in ("proj_matrix", "neighbor_topk", "clip_lerp_ndc", "specular_center_inverse"):
                O.set_option(name, 1)
                n = compare(f"oracle, {name}=1", oracle_render(O, scene, S, C, R, T), want)
                O.set_option(name, 0)
                if n < base:
                    print(f"  --> switch {name!r} removes {base - n} mismatch(es): pytorch3d takes the other choice here")
    # the pose arithmetic: look_at_view_transform / look_at_rotation against the oracle, with both trig variants
    import torch
    from pytorch3d.renderer import look_at_rotation, look_at_view_transform
    rng = np.random.default_rng(0)
    for trig in (0, 1):
        O.set_option("trig_fp32", trig)
        worst = 0.0
        for _ in range(200):
            r, el, az = 4.0, float(rng.uniform(-1.2, 1.2)), float(rng.uniform(-3, 3))
            C, R, T = O.pose_lookat(r, el, az)
            R3, T3 = look_at_view_transform(torch.tensor([r]), torch.tensor([el]), torch.tensor([az]), degrees=False)
            worst = max(worst, float(np.abs(R3[0].numpy() - R).max()), float(np.abs(T3[0].numpy() - T).max()))
            _, _, C2, R2, T2 = O.pose_step(np.zeros(2, np.float32), el, az, r)
            Rr = look_at_rotation(torch.tensor(C2[None]))[0].numpy()
            worst = max(worst, float(np.abs(Rr - R2).max()))
        print(f"pose (look_at_view_transform / look_at_rotation), trig_fp32={trig}: max |d| {worst:.3e}"
              f"{'  (bit-identical)' if worst == 0.0 else ''}")
    O.set_option("trig_fp32", 0)
    print("PINNED: every criterion met with the default switches" if failed_default == 0 else
          f"NOT PINNED: {failed_default} criteria failed with the default switches (see the switch lines above)")
    return 0 if failed_default == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
