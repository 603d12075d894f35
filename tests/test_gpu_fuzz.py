"""Seeded random sweep of the render against the oracle: scenes (teapot pair, box occluder, procedural 2-4 object
scenes, random triangle soups), camera poses (radius 0.6 .. 30, i.e. from inside the geometry -- faces cut or
removed at z_clip -- to far away, where hundreds of hits pile up on a few pixels), elevation up to the look-at pole,
faces_per_pixel 1 .. 100, image sizes 17 .. 72, tile shapes (compile-time 32x32 / 32x16, generic ones), with and
without back-face culling.  Every case compares the bit-exact outputs exactly and the tolerance outputs within the
tolerances of DESIGN.md section 3."""
import numpy as np
import pytest
import torch

from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.meshes import default_scene, pack_scene, procedural_scene

pytestmark = pytest.mark.gpu
RTOL = 1e-5
ATOL_A = 2e-6


def _soup(rng, n_obj, n_tri, s):
    """random triangle soup in view-friendly world space (objects around the origin, a few units across)"""
    objs = []
    for o in range(n_obj):
        ctr = rng.uniform(-1.0, 1.0, (n_tri, 1, 3)) + np.array([1.5 * o, 0.0, 0.0])
        size = 10 ** rng.uniform(-2.0, 0.0, (n_tri, 1, 1))
        tris = (ctr + rng.normal(size=(n_tri, 3, 3)) * size).astype(np.float32)
        verts = tris.reshape(-1, 3)
        faces = np.arange(len(verts), dtype=np.int32).reshape(-1, 3)
        objs.append((verts, faces))
    return pack_scene(objs)


def _case(rng, i):
    kind = ["teapot", "box", "proc", "soup"][i % 4]
    if kind in ("teapot", "box"):
        sc = default_scene(kind)
    elif kind == "proc":
        sc = procedural_scene(int(rng.integers(1000)), n_obj=int(rng.integers(2, 5)), subdiv=int(rng.integers(1, 4)))
    else:
        sc = _soup(rng, int(rng.integers(1, 4)), int(rng.integers(20, 200)), 1.0)
    S = int(rng.choice([17, 32, 40, 64, 72]))
    K = int(rng.choice([1, 3, 10, 50, 100]))
    radius = float(rng.choice([0.6, 1.0, 1.7, 2.3, 4.0, 4.0, 8.0, 30.0]))
    az = float(rng.uniform(-3.1, 3.1))
    el = float(rng.choice([0.0, rng.uniform(-1.2, 1.2), 1.5607, -1.5]))
    tile = [(0, 0), (0, 0), (32, 16), (16, 16), (24, 10), (64, 8)][int(rng.integers(6))]
    cull = bool(rng.integers(4) > 0)
    return sc, S, K, radius, az, el, tile, cull


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_renders_match_oracle(oracle, cuda_lib, seed):
    from occlusionenv_b200.engine import OcclusionEngine
    rng = np.random.default_rng(1234 + seed)
    n_cut = n_over = 0
    for i in range(24):
        sc, S, K, radius, az, el, (tw, th), cull = _case(rng, i + seed)
        tag = (seed, i, S, K, radius, round(az, 3), round(el, 3), tw, th, cull, sc.n_obj, int(sc.faces.shape[0]))
        cfg = RasterConfig(image_size=S, faces_per_pixel=K, cull_backfaces=cull, tile_w=tw, tile_h=th)
        eng = OcclusionEngine(sc, 1, cfg, debug_outputs=True)
        eng.reset(radius=radius, azimuth=az, elevation=el)
        st = int(eng.status[0])
        assert not (st & (1 | 4 | 8)), (tag, st)
        C, R, T = oracle.pose_lookat(radius, el, az)
        vproj = oracle.project(sc.verts, R, T)
        if not np.isfinite(vproj).all():
            continue  # a vertex exactly in the camera plane: inf / nan coordinates, nothing to compare
        alphas, nhits, straddle = [], [], False
        for o in range(sc.n_obj):
            v0, v1 = sc.obj_vert_start[o], sc.obj_vert_start[o + 1]
            f0, f1 = sc.obj_face_start[o], sc.obj_face_start[o + 1]
            fr = oracle.rasterize_clipped(vproj[v0:v1], sc.faces[f0:f1] - v0, S, oracle.BLUR_RADIUS, K, cull_backfaces=cull)
            alphas.append(oracle.silhouette(fr))
            nhits.append(fr.nhits)
            straddle |= fr.straddles
        scene = oracle.rasterize_clipped(vproj, sc.faces, S, 0.0, 1, cull_backfaces=cull)
        alphas, nhits = np.stack(alphas), np.stack(nhits)
        n_cut += int(straddle)
        n_over += int((nhits > K).any())
        assert bool(st & 16) == bool(straddle or scene.straddles), tag
        assert np.array_equal(eng.nhits[0].cpu().numpy(), nhits), tag
        assert np.array_equal(eng.pix_to_face[0].cpu().numpy(), scene.pix_to_face[..., 0]), tag
        assert np.array_equal(eng.obs[0, 3].cpu().numpy(), scene.zbuf[..., 0]), tag
        np.testing.assert_allclose(eng.alphas[0].cpu().numpy(), alphas, rtol=RTOL, atol=ATOL_A, err_msg=str(tag))
        occl = np.zeros((S, S), np.float32)
        for a in range(sc.n_obj):
            for b in range(a + 1, sc.n_obj):
                occl = occl + alphas[a] * alphas[b]
        np.testing.assert_allclose(eng.occl[0].cpu().numpy(), occl, rtol=RTOL, atol=4e-6, err_msg=str(tag))
        loss = float(np.sum(occl.astype(np.float64) ** 2))
        np.testing.assert_allclose(float(eng.loss[0]), loss, rtol=2e-5, atol=1e-5, err_msg=str(tag))
    assert n_over >= 3, "the sweep must exercise the nearest-K rule"
    assert n_cut >= 1, "the sweep must exercise faces cut at z_clip"


def test_random_gradients_match_dense_autograd(oracle, cuda_lib):
    """Differentiable step on random poses / image sizes / K (the nearest-K rule with tangents included): d reward /
    d action within 1e-3 of the float64 autograd of the dense formulation (oracle/dense_torch.py)."""
    from occlusionenv_b200.engine import OcclusionEngine
    from oracle import dense_torch as D
    import math
    rng = np.random.default_rng(99)
    checked = 0
    for i in range(10):
        sc = default_scene(["box", "teapot"][i % 2])
        S = int(rng.choice([32, 48, 64]))
        K = int(rng.choice([5, 20, 100]))
        radius = float(rng.choice([4.0, 5.0, 6.0]))  # no face within z_clip: no gradient flows through cut faces
        az0 = float(math.pi / 2 + rng.uniform(-0.5, 0.5))
        el0 = float(rng.uniform(-0.3, 0.3))
        act = rng.normal(size=2).astype(np.float32)
        tag = (i, S, K, radius, round(az0, 3), round(el0, 3), act.tolist())
        eng = OcclusionEngine(sc, 1, RasterConfig(image_size=S, faces_per_pixel=K))
        eng.reset(radius=radius, azimuth=az0, elevation=el0)
        prev, mass = float(eng.full_reward[0]), float(eng.object_mass[0])
        eng.step(torch.tensor(act[None], device="cuda"), with_grad=True)
        eng.check_status()
        r, loss, g, _ = D.reward_and_grad(sc, S, act.astype(np.float64), el0, az0, radius, prev, mass, float(oracle.PROJ_SCALE),
                                          float(oracle.BLUR_RADIUS), float(oracle.SIGMA), K=K)
        np.testing.assert_allclose(float(eng.loss[0]), loss, rtol=2e-5, atol=1e-5, err_msg=str(tag))
        g_gpu = eng.grad_action[0].cpu().numpy()
        scale = np.abs(g).max()
        if scale < 1e-7:
            continue  # no overlap at this pose: nothing to compare
        assert np.abs(g_gpu - g).max() <= 1e-3 * scale + 1e-7, (tag, g_gpu, g)
        checked += 1
    assert checked >= 6
