"""GPU parity of the whole transition (reset + step, reward/done/state, gradient, golden fixtures) and of
the reference-facing API (OcclusionEnv, BatchedOcclusionVecEnv, SimpleVecEnv)."""
import os

import numpy as np
import pytest
import torch

from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.meshes import default_scene, procedural_scene

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-5
ATOL_A = 2e-6  # 16 ulp of 1.0: alpha = 1 - prod_{k<=100}(1-p_k) in fp32 (see test_gpu_render_parity.py)


def _engine(sc, n, S, **kw):
    from occlusionenv_b200.engine import OcclusionEngine
    return OcclusionEngine(sc, n, RasterConfig(image_size=S), **kw)


def test_pose_kernels_bit_identical_to_oracle(oracle, cuda_lib):
    sc = default_scene("box")
    rng = np.random.default_rng(0)
    n = 257
    az = rng.uniform(-3, 3, n).astype(np.float32)
    el = rng.uniform(-1.2, 1.2, n).astype(np.float32)
    act = rng.normal(size=(n, 2)).astype(np.float32)
    act[3] = 0.0
    eng = _engine(sc, n, 16)
    eng.set_pose(4.0, torch.tensor(az), torch.tensor(el))
    from occlusionenv_b200 import _lib as L
    import ctypes
    cam = torch.zeros(n, L.OCCL_CAM_STRIDE, device="cuda")
    L.check(cuda_lib.occl_pose_lookat(ctypes.byref(eng.c), n, eng.c_state, cam.data_ptr(), None), "lookat")
    cam_l = cam.cpu().numpy()
    L.check(cuda_lib.occl_pose_step(ctypes.byref(eng.c), n, torch.tensor(act, device="cuda").data_ptr(), eng.c_state,
                                    cam.data_ptr(), None), "step")
    cam_s = cam.cpu().numpy()
    el2, az2 = eng.elevation.cpu().numpy(), eng.azimuth.cpu().numpy()
    for e in range(n):
        C, R, T = oracle.pose_lookat(4.0, el[e], az[e])
        assert np.array_equal(cam_l[e, :9].reshape(3, 3), R) and np.array_equal(cam_l[e, 9:12], T) and np.array_equal(cam_l[e, 12:15], C)
        oe, oa, C, R, T = oracle.pose_step(act[e], el[e], az[e], 4.0)
        assert oe == el2[e] and oa == az2[e]
        assert np.array_equal(cam_s[e, :9].reshape(3, 3), R) and np.array_equal(cam_s[e, 9:12], T) and np.array_equal(cam_s[e, 12:15], C)


@pytest.mark.parametrize("occ", ["teapot", "box"])
def test_golden_fixtures(cuda_lib, occ):
    g = np.load(os.path.join(GOLD, f"scene_{occ}_128.npz"))
    sc = default_scene(occ)
    n = len(g["poses"])
    eng = _engine(sc, n, 128, debug_outputs=True)
    R = torch.tensor(np.stack([g[f"R{k}"] for k in range(n)]), device="cuda").contiguous()
    T = torch.tensor(np.stack([g[f"T{k}"] for k in range(n)]), device="cuda").contiguous()
    C = torch.tensor(np.stack([g[f"C{k}"] for k in range(n)]), device="cuda").contiguous()
    eng.render(R, T, C)
    eng.check_status()
    for k in range(n):
        assert np.array_equal(eng.pix_to_face[k].cpu().numpy().astype(np.int16), g[f"pix_to_face{k}"])
        assert np.array_equal(eng.obs[k, 3].cpu().numpy(), g[f"zbuf{k}"])
        assert np.array_equal(eng.n_covered[k].cpu().numpy(), g[f"n_covered{k}"])
        assert np.array_equal(eng.n_visible[k].cpu().numpy(), g[f"n_visible{k}"])
        assert int(eng.nhits[k].max()) == int(g[f"nhits_max{k}"])
        np.testing.assert_allclose(eng.alphas[k].cpu().numpy(), g[f"alphas{k}"], rtol=RTOL, atol=ATOL_A)
        np.testing.assert_allclose(eng.obs[k, 0].cpu().numpy(), g[f"rgb{k}"], rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(float(eng.loss[k]), float(g[f"loss{k}"]), rtol=RTOL, atol=1e-6)


def test_golden_fixture_near_camera_cut_faces(cuda_lib):
    """tests/golden/scene_teapot_near_64.npz: camera inside the target's bounding box -- faces cut at z_clip
    (clip_faces cases 3/4) are visible on a few hundred pixels, faces nearer than z_clip are removed."""
    g = np.load(os.path.join(GOLD, "scene_teapot_near_64.npz"))
    sc = default_scene("teapot")
    eng = _engine(sc, 1, 64, debug_outputs=True)
    R, T, C = (torch.tensor(g[k][None], device="cuda").contiguous() for k in ("R", "T", "C"))
    eng.render(R, T, C)
    assert eng.check_status() & 16, "the clip-capable kernel must have taken this env"
    assert np.array_equal(eng.pix_to_face[0].cpu().numpy().astype(np.int16), g["pix_to_face"])
    assert np.array_equal(eng.obs[0, 3].cpu().numpy(), g["zbuf"])
    assert np.array_equal(eng.nhits[0].cpu().numpy().astype(np.int16), g["nhits"])
    assert np.array_equal(eng.n_covered[0].cpu().numpy(), g["n_covered"])
    assert np.array_equal(eng.n_visible[0].cpu().numpy(), g["n_visible"])
    np.testing.assert_allclose(eng.alphas[0].cpu().numpy(), g["alphas"], rtol=RTOL, atol=ATOL_A)
    np.testing.assert_allclose(eng.obs[0, 0].cpu().numpy(), g["rgb"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(eng.bary[0].cpu().numpy(), g["bary"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(eng.loss[0]), float(g["loss"]), rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("occ", ["teapot", "box"])
def test_pose_gradient_golden_fixture(cuda_lib, occ):
    """d loss / d (el, az) of the differentiable step against tests/golden/grad_*_128.npz (float64 autograd), 1e-3."""
    g = np.load(os.path.join(GOLD, f"grad_{occ}_128.npz"))
    poses, want = g["poses"], g["dloss_del_daz"]
    n = len(poses)
    eng = _engine(default_scene(occ), n, 128)
    eng.reset(radius=4.0, azimuth=torch.tensor(poses[:, 0].copy()), elevation=torch.tensor(poses[:, 1].copy()))
    mass = eng.object_mass.cpu().numpy().astype(np.float64)
    eng.step(torch.zeros(n, 2, device="cuda"), with_grad=True)   # zero action: the pose stays, grad_action = step * d reward / d (el, az)
    eng.check_status()
    got = -eng.grad_action.cpu().numpy().astype(np.float64) * mass[:, None] / 0.05
    for k in range(n):
        scale = np.abs(want[k]).max()
        if scale < 1e-3:
            assert np.abs(got[k]).max() < 1e-3, (k, got[k])
        else:
            assert np.abs(got[k] - want[k]).max() <= 1e-3 * scale, (k, got[k], want[k])


def test_trajectory_matches_oracle_state_machine(oracle, cuda_lib):
    """reset + several steps for a batch of envs: reward / done / loss / state, every step."""
    sc = default_scene("teapot")
    S, n, steps = 64, 6, 3
    rng = np.random.default_rng(1)
    az0 = rng.uniform(np.pi / 2 - 0.6, np.pi / 2 + 0.6, n).astype(np.float32)
    az0[0] = 0.0  # no occlusion -> done on the first step
    el0 = rng.uniform(-0.3, 0.3, n).astype(np.float32)
    acts = rng.normal(size=(steps, n, 2)).astype(np.float32)
    acts[1, 2] = 0.0
    eng = _engine(sc, n, S, debug_outputs=True)
    eng.reset(radius=4.0, azimuth=torch.tensor(az0), elevation=torch.tensor(el0))
    refs = []
    for e in range(n):
        r = oracle.OracleOcclusionEnv(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, img_size=S)
        obs = r.reset(radius=4.0, azimuth=az0[e], elevation=el0[e])
        np.testing.assert_allclose(eng.obs[e].cpu().numpy(), obs[0], rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(float(eng.full_reward[e]), float(r.fullReward), rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(float(eng.object_mass[e]), float(r.objectMass), rtol=RTOL)
        refs.append(r)
    for t in range(steps):
        eng.step(torch.tensor(acts[t], device="cuda"))
        eng.check_status()
        for e, r in enumerate(refs):
            obs, rew, done, info = r.step(acts[t, e])
            assert np.array_equal(eng.pix_to_face[e].cpu().numpy(), r.last.pix_to_face)
            assert np.array_equal(eng.n_visible[e].cpu().numpy(), r.last.n_visible)
            np.testing.assert_allclose(eng.obs[e].cpu().numpy(), obs[0], rtol=RTOL, atol=1e-6)
            np.testing.assert_allclose(eng.occl[e].cpu().numpy(), info["full_state"], rtol=2 * RTOL, atol=2 * ATOL_A)
            np.testing.assert_allclose(float(eng.reward[e]), float(rew), rtol=RTOL, atol=2e-6)
            np.testing.assert_allclose(float(eng.loss[e]), float(info["full_reward"]), rtol=RTOL, atol=1e-6)
            assert bool(eng.done[e]) == bool(done)
            np.testing.assert_array_equal(eng.position[e].cpu().numpy(), info["position"])
            assert float(eng.elevation[e]) == float(r.elevation) and float(eng.azimuth[e]) == float(r.azimuth)
    assert bool(eng.done[0])


@pytest.mark.parametrize("occ", ["teapot", "box"])
def test_gradient_to_action_matches_dense_autograd(oracle, cuda_lib, occ):
    """north-star tolerance: pose gradients within 1e-3 relative (vs float64 autograd of oracle (b))."""
    from oracle import dense_torch as D
    sc = default_scene(occ)
    S = 64
    cases = [(1.45, 0.1, (0.3, -1.0)), (1.7, -0.2, (-0.5, 0.2)), (1.3, 0.25, (1.0, 1.0)), (1.55, 0.0, (0.0, 0.0))]
    n = len(cases)
    eng = _engine(sc, n, S)
    az0 = np.array([c[0] for c in cases], np.float32)
    el0 = np.array([c[1] for c in cases], np.float32)
    eng.reset(radius=4.0, azimuth=torch.tensor(az0), elevation=torch.tensor(el0))
    prev = eng.full_reward.cpu().numpy().copy()
    mass = eng.object_mass.cpu().numpy().copy()
    act = np.array([c[2] for c in cases], np.float32)
    eng.step(torch.tensor(act, device="cuda"), with_grad=True)
    eng.check_status()
    g_gpu = eng.grad_action.cpu().numpy()
    for e in range(n):
        r, loss, g, _ = D.reward_and_grad(sc, S, act[e].astype(np.float64), el0[e], az0[e], 4.0, float(prev[e]), float(mass[e]),
                                          float(oracle.PROJ_SCALE), float(oracle.BLUR_RADIUS), float(oracle.SIGMA))
        np.testing.assert_allclose(float(eng.loss[e]), loss, rtol=1e-5)
        scale = max(np.abs(g).max(), 1e-6)
        assert np.abs(g_gpu[e] - g).max() <= 1e-3 * scale, (e, g_gpu[e], g)
    # the non-differentiable forward gives the same reward
    eng2 = _engine(sc, n, S)
    eng2.reset(radius=4.0, azimuth=torch.tensor(az0), elevation=torch.tensor(el0))
    eng2.step(torch.tensor(act, device="cuda"), with_grad=False)
    np.testing.assert_allclose(eng2.reward.cpu().numpy(), eng.reward.cpu().numpy(), rtol=1e-6, atol=1e-7)


def test_occlusion_env_dropin_contract(oracle, cuda_lib):
    """Shapes / dtypes / attributes of environment.py:201-402 and autograd through reward (demo.py:85-91)."""
    from occlusionenv_b200.environment import OcclusionEnv
    S = 64
    env = OcclusionEnv(img_size=S)
    assert env.observation_space.shape == (4, S, S) and env.action_space.shape == (2,)
    assert env.step_size == 0.05 and env.normWithObjectSize is False and env.renderMode == ""
    env.seed(3)
    obs = env.reset(azimuth=1.5)
    assert obs.shape == (1, 4, S, S) and obs.dtype == torch.float32 and obs.is_cuda
    assert len(env.meshes) == 3 and float(env.camera_position.abs().sum()) == 0.0
    action = torch.nn.Parameter(torch.tensor([0.3, -1.0]))
    obs, reward, finished, info = env.step(action)
    assert obs.shape == (1, 4, S, S) and reward.dim() == 0 and finished.dtype == torch.bool and finished.dim() == 0
    assert info["full_state"].shape == (1, S, S, 4) and info["position"].shape == (3,) and info["full_reward"].dim() == 0
    reward.backward()
    assert action.grad is not None and action.grad.shape == (2,) and torch.isfinite(action.grad).all()
    assert float(action.grad.abs().sum()) > 0
    # same numbers as the oracle env
    sc = default_scene("teapot")
    ref = oracle.OracleOcclusionEnv(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, img_size=S)
    ref.reset(azimuth=1.5)
    _, r, d, _ = ref.step(np.array([0.3, -1.0], np.float32))
    np.testing.assert_allclose(float(reward), float(r), rtol=RTOL, atol=2e-6)
    assert bool(finished) == bool(d)
    rgba, depth = env.render()
    assert rgba.shape == (1, S, S, 4) and depth.shape == (1, S, S, 1)
    np.testing.assert_allclose(rgba[0, ..., 0].cpu().numpy(), obs[0, 0].detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
    # numpy actions and no-grad path
    o2, r2, f2, _ = env.step(np.array([0.0, 0.0], np.float32))
    assert abs(float(r2) + 0.2) < 1e-6 or bool(f2)
    env.close()


def test_batched_vecenv_matches_sequential_reference_loop(oracle, cuda_lib):
    """BatchedOcclusionVecEnv == the reference's SimpleVecEnv loop (SubProcVecEnv.py:203-220) incl. auto-reset."""
    from occlusionenv_b200.SubProcVecEnv import BatchedOcclusionVecEnv
    sc = default_scene("teapot")
    S, n = 32, 5
    venv = BatchedOcclusionVecEnv(n, data=None, img_size=S, keep_terminal_obs=True)
    az0 = np.array([1.5, 0.0, 1.3, 1.8, 0.05], np.float32)  # envs 1 and 4 start (almost) unoccluded
    obs = venv.reset(azimuth=torch.tensor(az0))
    assert obs.shape == (n, 4, S, S)
    refs = [oracle.OracleOcclusionEnv(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, img_size=S) for _ in range(n)]
    ovec = oracle.OracleSimpleVecEnv(refs)
    ovec.reset(azimuths=az0)
    rng = np.random.default_rng(5)
    for t in range(3):
        a = rng.normal(size=(n, 2)).astype(np.float32)
        obs, rews, dones, infos = venv.step(torch.tensor(a))
        o_obs, o_rews, o_dones, o_infos = ovec.step(a)
        assert obs.shape == (n, 4, S, S) and rews.shape == (n,) and dones.shape == (n,) and dones.dtype == torch.bool
        assert len(infos) == n
        assert np.array_equal(dones.cpu().numpy(), o_dones)
        np.testing.assert_allclose(rews.cpu().numpy(), o_rews, rtol=RTOL, atol=2e-6)
        np.testing.assert_allclose(obs.cpu().numpy(), o_obs, rtol=RTOL, atol=1e-6)
        for e in range(n):
            if o_dones[e]:
                np.testing.assert_allclose(infos[e]["terminal_observation"][0].cpu().numpy(),
                                           o_infos[e]["terminal_observation"][0], rtol=RTOL, atol=1e-6)
                np.testing.assert_allclose(float(venv.engine.full_reward[e]), float(refs[e].fullReward), rtol=RTOL, atol=1e-6)
        assert o_dones.any() or t > 0
    # differentiable batched step (train_predict.py:48-52)
    step = torch.nn.Parameter(torch.randn(n, 2, device="cuda"))
    obs, rews, dones, infos = venv.step(step)
    rews.sum().backward()
    assert step.grad.shape == (n, 2) and torch.isfinite(step.grad).all()


def test_simple_vecenv_reference_shapes(cuda_lib):
    from occlusionenv_b200.SubProcVecEnv import SimpleVecEnv
    from occlusionenv_b200.environment import OcclusionEnv
    venv = SimpleVecEnv([lambda: OcclusionEnv(img_size=32) for _ in range(2)])
    obs = venv.reset()
    assert obs.shape == (2, 1, 4, 32, 32)  # reference quirk B-8
    obs, rews, dones, infos = venv.step(torch.randn(2, 2))
    assert obs.shape == (2, 4, 32, 32) and rews.shape == (2,) and dones.shape == (2,) and len(infos) == 2
    assert venv.get_attr("step_size") == [0.05, 0.05]
    venv.set_attr("step_size", 0.1, indices=0)
    assert venv.get_attr("step_size") == [0.1, 0.05]
    assert len(venv.get_images()) == 2
    venv.close()


def test_three_object_per_env_meshes(oracle, cuda_lib):
    """Config-3 shape at test size: per-env procedural meshes, 3 objects in the ShapeNet layout; K=100 cut
    is the common case here."""
    from occlusionenv_b200.engine import OcclusionEngine
    S, n = 64, 2
    scenes = [procedural_scene(s, n_obj=3, subdiv=3) for s in (11, 12)]
    eng = OcclusionEngine(None, n, RasterConfig(image_size=S), debug_outputs=True, per_env_scenes=scenes)
    az0 = np.array([0.4, -0.3], np.float32)
    eng.reset(radius=4.0, azimuth=torch.tensor(az0), elevation=0.1)
    act = np.array([[1.0, 0.5], [-0.3, 0.9]], np.float32)
    eng.step(torch.tensor(act, device="cuda"))
    st = eng.status.cpu().numpy()
    assert not (st & (4 | 8)).any(), st
    for e, sc in enumerate(scenes):
        ref = oracle.OracleOcclusionEnv(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, img_size=S)
        ref.reset(radius=4.0, azimuth=az0[e], elevation=0.1)
        _, rew, done, _ = ref.step(act[e])
        assert not (st[e] & 1)
        assert np.array_equal(eng.pix_to_face[e].cpu().numpy(), ref.last.pix_to_face)
        assert np.array_equal(eng.nhits[e].cpu().numpy(), ref.last.nhits)
        np.testing.assert_allclose(eng.alphas[e].cpu().numpy(), ref.last.alphas, rtol=RTOL, atol=ATOL_A)
        np.testing.assert_allclose(float(eng.reward[e]), float(rew), rtol=RTOL, atol=2e-6)


def test_tile_shapes_and_odd_image_size(oracle, cuda_lib):
    """Ragged case: image size not a multiple of the tile; several tile shapes give identical results."""
    from occlusionenv_b200.engine import OcclusionEngine
    sc = default_scene("box")
    S = 50
    _, _, C, R, T = oracle.pose_step(np.zeros(2, np.float32), 0.1, 1.4, 4.0)
    ref = oracle.render_scene(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, S, C, R, T)
    Rt, Tt, Ct = (torch.tensor(x[None], device="cuda").contiguous() for x in (R, T, C))
    for tw, th in [(0, 0), (32, 16), (16, 16), (50, 8), (64, 32), (43, 43), (7, 5)]:  # (0,0) and (32,16): the compile-time tiles
        eng = OcclusionEngine(sc, 1, RasterConfig(image_size=S, tile_w=tw, tile_h=th), debug_outputs=True)
        eng.render(Rt, Tt, Ct)
        eng.check_status()
        assert np.array_equal(eng.pix_to_face[0].cpu().numpy(), ref.pix_to_face), (tw, th)
        assert np.array_equal(eng.nhits[0].cpu().numpy(), ref.nhits)
        np.testing.assert_allclose(eng.alphas[0].cpu().numpy(), ref.alphas, rtol=RTOL, atol=ATOL_A)
        np.testing.assert_allclose(eng.obs[0].cpu().numpy(), ref.obs, rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("occ", ["box", "teapot"])
def test_fast_path_equals_exact_path_at_full_batch(cuda_lib, occ):
    """Size-independent property at the BASELINE batch shape: the guarded fast pair evaluation takes
    exactly the decisions of the reference-order arithmetic (debug_exact) for every pixel of every env."""
    from occlusionenv_b200.engine import OcclusionEngine
    import math
    sc = default_scene(occ)
    N, S = 2048, 128
    g = torch.Generator().manual_seed(0)
    az = (math.pi / 2 - 0.6) + 1.2 * torch.rand(N, generator=g)
    el = -0.3 + 0.6 * torch.rand(N, generator=g)
    act = torch.randn(N, 2, generator=g).cuda()
    res = []
    for exact in (False, True):
        eng = OcclusionEngine(sc, N, RasterConfig(image_size=S, debug_exact=exact), debug_outputs=True)
        eng.reset(radius=4.0, azimuth=az, elevation=el)
        eng.step(act, with_grad=True)
        eng.check_status()
        res.append(eng)
    a, b = res
    assert torch.equal(a.pix_to_face, b.pix_to_face)
    assert torch.equal(a.nhits, b.nhits)
    assert torch.equal(a.n_covered, b.n_covered) and torch.equal(a.n_visible, b.n_visible)
    assert torch.equal(a.obs, b.obs), "observation (depth + shading) does not depend on the fast path"
    assert torch.equal(a.done, b.done)
    torch.testing.assert_close(a.alphas, b.alphas, rtol=RTOL, atol=ATOL_A)
    torch.testing.assert_close(a.loss, b.loss, rtol=RTOL, atol=1e-6)
    torch.testing.assert_close(a.reward, b.reward, rtol=RTOL, atol=2e-6)
    ga, gb = a.grad_action, b.grad_action
    # 1e-3 relative to the env's gradient, with an absolute floor for envs whose overlap (and gradient) is ~0:
    # the per-pixel tangent sums are fp32 atomics whose order differs from run to run
    scale = gb.abs().max(dim=1, keepdim=True).values
    assert bool(((ga - gb).abs() <= 1e-3 * scale + 1e-6).all())
    assert int(a.nhits.max()) > 100  # the K=100 cut is exercised


def test_hoisted_reciprocal_division_is_ieee_exact(cuda_lib):
    """The rasteriser divides by per-face constants through a hoisted correctly-rounded reciprocal and
    two Markstein steps; over 4e9 random operand pairs of its guarded domain every quotient is
    bit-identical to IEEE `a / b`."""
    from occlusionenv_b200 import _lib as L
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for seed in (1, 2):
        L.check(cuda_lib.occl_selftest_div(2_000_000_000, seed, bad.data_ptr(), None), "occl_selftest_div")
        torch.cuda.synchronize()
        assert int(bad.item()) == 0


@pytest.mark.parametrize("seed,az,el", [(21, 0.5, 0.15), (4, -0.5, 0.1)])
def test_dense_meshes_many_overflowing_pixels(oracle, cuda_lib, seed, az, el):
    """Config-3 meshes (three 20480-face objects, ShapeNet layout) at 64^2: half of the silhouette pixels
    collect more than faces_per_pixel=100 hits (700 to 1000 at the most crowded ones): every path of the
    nearest-K rule is exercised (strong-hit shortcut, batched selection, two-scan threshold selection, several
    overflow rounds per tile)."""
    from occlusionenv_b200.engine import OcclusionEngine
    S = 64
    sc = procedural_scene(seed, n_obj=3, subdiv=5)
    eng = OcclusionEngine(sc, 1, RasterConfig(image_size=S), debug_outputs=True)
    _, _, C, R, T = oracle.pose_step(np.zeros(2, np.float32), el, az, 4.0)
    Rt, Tt, Ct = (torch.tensor(x[None], device="cuda").contiguous() for x in (R, T, C))
    eng.render(Rt, Tt, Ct)
    st = int(eng.status[0])
    assert not (st & (1 | 4)), st
    ref = oracle.render_scene(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, S, C, R, T)
    assert (ref.nhits > 100).sum() > 150 and ref.nhits.max() > 500, "test scene must overflow massively"
    assert np.array_equal(eng.nhits[0].cpu().numpy(), ref.nhits)
    assert np.array_equal(eng.pix_to_face[0].cpu().numpy(), ref.pix_to_face)
    np.testing.assert_allclose(eng.alphas[0].cpu().numpy(), ref.alphas, rtol=RTOL, atol=ATOL_A)
    np.testing.assert_allclose(float(eng.loss[0]), float(ref.loss), rtol=RTOL, atol=1e-6)
    # differentiable variant takes the same decisions
    eng2 = OcclusionEngine(sc, 1, RasterConfig(image_size=S, debug_exact=True), debug_outputs=True)
    eng2.render(Rt, Tt, Ct)
    assert torch.equal(eng.nhits, eng2.nhits) and torch.equal(eng.pix_to_face, eng2.pix_to_face)
    torch.testing.assert_close(eng.alphas, eng2.alphas, rtol=RTOL, atol=ATOL_A)


@pytest.mark.parametrize("variant", ["far_camera", "k10", "no_cull", "norm_object_size", "tiny_image", "near_camera_clip", "near_camera_clip_k3"])
def test_configuration_variants_match_oracle(oracle, cuda_lib, variant):
    """Edge cases of the raster configuration: sub-pixel faces (all hits pile up on a few pixels), a small K
    (the nearest-K rule everywhere), no back-face culling (negative-area faces take the reference-order path),
    the normWithObjectSize reward branch (environment.py:324), a 16x16 image, and a camera inside the occluder's
    bounding box (faces cut at z_clip = znear/2 by clip_faces; the differentiable step refuses such frames)."""
    from occlusionenv_b200.engine import OcclusionEngine
    sc = default_scene("teapot")
    S, K, cull, radius, norm = 64, 100, True, 4.0, False
    if variant == "far_camera":
        radius = 30.0
    elif variant == "k10":
        K = 10
    elif variant == "no_cull":
        cull = False
    elif variant == "norm_object_size":
        norm = True
    elif variant == "tiny_image":
        S = 16
    elif variant.startswith("near_camera_clip"):
        radius = 1.0  # camera inside the target teapot's bounding box: clip_faces removes faces and cuts visible ones
        if variant.endswith("k3"):
            K = 3     # ... and the nearest-K rule has to pick among hits on cut faces
    cfg = RasterConfig(image_size=S, faces_per_pixel=K, cull_backfaces=cull, norm_with_object_size=norm)
    eng = OcclusionEngine(sc, 1, cfg, debug_outputs=True)
    az, el = (1.0, 0.1) if variant.startswith("near_camera_clip") else (1.45, 0.1)
    eng.reset(radius=radius, azimuth=az, elevation=el)
    st = int(eng.status[0])
    assert not (st & (1 | 4)), st
    C, R, T = oracle.pose_lookat(radius, el, az)
    vproj = oracle.project(sc.verts, R, T)
    alphas, nhits = [], []
    for i in range(sc.n_obj):
        v0, v1 = sc.obj_vert_start[i], sc.obj_vert_start[i + 1]
        f0, f1 = sc.obj_face_start[i], sc.obj_face_start[i + 1]
        fr = oracle.rasterize_clipped(vproj[v0:v1], sc.faces[f0:f1] - v0, S, oracle.BLUR_RADIUS, K, cull_backfaces=cull)
        alphas.append(oracle.silhouette(fr))
        nhits.append(fr.nhits)
        if variant.startswith("near_camera_clip") and i == 0:
            assert fr.straddles, "the test pose must cut faces of the target"
            zf = vproj[v0:v1][sc.faces[f0:f1] - v0][:, :, 2]
            cut = np.nonzero(((zf < 0.5).sum(1) % 3) != 0)[0]
            seen = np.isin(fr.pix_to_face, cut) & (fr.pix_to_face >= 0)
            assert seen.any(-1).sum() > 100, "cut faces must be visible"
            if K == 3:
                assert (seen.any(-1) & (fr.nhits > K)).sum() > 10, "cut faces must take part in the nearest-K rule"
    scene = oracle.rasterize_clipped(vproj, sc.faces, S, 0.0, 1, cull_backfaces=cull)
    assert np.array_equal(eng.nhits[0].cpu().numpy(), np.stack(nhits))
    assert np.array_equal(eng.pix_to_face[0].cpu().numpy(), scene.pix_to_face[..., 0])
    assert np.array_equal(eng.obs[0, 3].cpu().numpy(), scene.zbuf[..., 0])
    np.testing.assert_allclose(eng.alphas[0].cpu().numpy(), np.stack(alphas), rtol=RTOL, atol=ATOL_A)
    occl = alphas[0] * alphas[1]
    loss = float(np.sum(occl.astype(np.float64) ** 2))
    np.testing.assert_allclose(float(eng.loss[0]), loss, rtol=RTOL, atol=1e-6)
    objsq = float(np.sum((alphas[0] + alphas[1]).astype(np.float64) ** 2))
    mass = (objsq if norm else loss) + 1.0
    np.testing.assert_allclose(float(eng.object_mass[0]), mass, rtol=RTOL)
    if variant in ("far_camera", "k10", "near_camera_clip_k3"):
        assert (np.stack(nhits) > K).any()
    if variant.startswith("near_camera_clip"):
        np.testing.assert_allclose(eng.bary[0].cpu().numpy(), scene.bary[..., 0, :], rtol=1e-5, atol=1e-6)
        # the differentiable step takes cut faces too (test_gradient_through_faces_cut_at_z_clip checks the values)
        eng.step(torch.zeros(1, 2, device="cuda"), with_grad=True)
        assert not (int(eng.status[0]) & 1) and (int(eng.status[0]) & 16)
        eng.check_status()
        assert torch.isfinite(eng.grad_action).all()


def test_gradient_through_faces_cut_at_z_clip(oracle, cuda_lib):
    """Camera so close that faces straddle z_clip = znear / 2 (clip_faces cases 3 / 4) AND the objects occlude each
    other: d reward / d action through the cut triangles -- their plane intersections move with x, y and z of the
    uncut vertices -- against float64 autograd of the clip-aware dense formulation, 1e-3 relative.  The discrete
    decisions (hit sets, the neighbour rule on the shared diagonal of a cut quadrilateral, nearest K) are the fp32
    oracle's in both, see oracle/dense_torch.py::frozen_hit_masks."""
    from oracle import dense_torch as D
    sc = default_scene("teapot")
    S = 48
    cases = [(1.8, 1.5, 0.3, (0.3, -0.4)), (1.8, 1.8, 0.3, (-1.0, 0.2)), (2.2, 1.5, 0.3, (0.5, 0.5)), (2.2, 1.8, 0.0, (0.0, 1.0))]
    n = len(cases)
    eng = _engine(sc, n, S)
    eng.set_pose(torch.tensor([c[0] for c in cases]), torch.tensor([c[1] for c in cases]), torch.tensor([c[2] for c in cases]))
    eng.full_reward.zero_()
    eng.object_mass.fill_(1.0)
    act = np.array([c[3] for c in cases], np.float32)
    eng.step(torch.tensor(act, device="cuda"), with_grad=True)
    st = eng.status.cpu().numpy()
    assert (st & 16).all(), "every case must cut faces"
    assert not (st & (1 | 4 | 8)).any()
    eng.check_status()
    g_gpu = eng.grad_action.cpu().numpy()
    for e, (r, az, el, a) in enumerate(cases):
        _, loss, g, _ = D.reward_and_grad(sc, S, np.asarray(a, np.float64), el, az, r, 0.0, 1.0, float(oracle.PROJ_SCALE),
                                          float(oracle.BLUR_RADIUS), float(oracle.SIGMA), freeze_hits=True)
        assert loss > 1.0, "the case must show occlusion"
        np.testing.assert_allclose(float(eng.loss[e]), loss, rtol=2e-5)
        scale = max(np.abs(g).max(), 1e-6)
        assert np.abs(g_gpu[e] - g).max() <= 1e-3 * scale, (e, g_gpu[e], g)


def test_gradient_with_recorded_overflow_slots(oracle, cuda_lib):
    """Differentiable step on the 32x16 compile-time tile (three objects) with K = 10, so that most covered slots exceed
    K: the evaluate-once path (hits recorded in the main phase, selection from the slab, tangents of the kept hits
    re-evaluated by face index) against float64 autograd through the fp32 oracle's nearest-K sets."""
    from occlusionenv_b200.engine import OcclusionEngine
    from oracle import dense_torch as D
    S, K = 64, 10
    scenes = [procedural_scene(s, n_obj=3, subdiv=3) for s in (2, 3)]
    n = len(scenes)
    eng = OcclusionEngine(None, n, RasterConfig(image_size=S, faces_per_pixel=K), per_env_scenes=scenes)
    assert (int(eng.c.tile_w), int(eng.c.tile_h)) == (32, 16)
    az0, el0 = np.array([-0.35, 0.3], np.float32), np.array([0.1, 0.1], np.float32)
    eng.set_pose(4.0, torch.tensor(az0), torch.tensor(el0))
    eng.full_reward.zero_()
    eng.object_mass.fill_(1.0)
    act = np.array([[1.0, 0.5], [-0.3, 0.9]], np.float32)
    eng.step(torch.tensor(act, device="cuda"), with_grad=True)
    assert eng.check_status() & 2, "K must be live"
    g_gpu = eng.grad_action.cpu().numpy()
    for e, sc in enumerate(scenes):
        _, loss, g, _ = D.reward_and_grad(sc, S, act[e].astype(np.float64), float(el0[e]), float(az0[e]), 4.0, 0.0, 1.0,
                                          float(oracle.PROJ_SCALE), float(oracle.BLUR_RADIUS), float(oracle.SIGMA), K=K,
                                          freeze_hits=True)
        assert loss > 1.0
        np.testing.assert_allclose(float(eng.loss[e]), loss, rtol=2e-5)
        scale = max(np.abs(g).max(), 1e-6)
        assert np.abs(g_gpu[e] - g).max() <= 1e-3 * scale, (e, g_gpu[e], g)


def test_adversarial_triangles_match_oracle(oracle, cuda_lib):
    """Hand-made nasty geometry through occl_render with an identity camera: slivers with one edge shorter
    than 1e-4 (degenerate-edge branch of the point-segment distance), needle triangles, exact duplicates
    (depth ties -> lower face index), faces far outside the image, sub-pixel faces, pixel-centre-aligned
    vertices, faces with |ndc| > 4 and clockwise faces (culled).  Every guard of the fast path has to hand these
    to the reference-order routine or decide them identically."""
    from occlusionenv_b200.engine import OcclusionEngine
    from occlusionenv_b200.meshes import pack_scene
    rng = np.random.default_rng(7)
    S = 64
    s = float(oracle.PROJ_SCALE)
    tris = []

    def add(p0, p1, p2, z=(2.0, 2.0, 2.0)):
        # p* are NDC xy; convert to view/world coordinates (identity camera): x_ndc = s x / z
        tris.append([[p0[0] * z[0] / s, p0[1] * z[0] / s, z[0]], [p1[0] * z[1] / s, p1[1] * z[1] / s, z[1]],
                     [p2[0] * z[2] / s, p2[1] * z[2] / s, z[2]]])

    c = lambda i: -1.0 + (2 * i + 1) / S  # pixel centre
    for _ in range(60):  # random small / medium faces, random depth per vertex
        ctr = rng.uniform(-0.9, 0.9, 2)
        r = 10 ** rng.uniform(-3, -0.5)
        pts = ctr + rng.normal(size=(3, 2)) * r
        add(*pts, z=tuple(rng.uniform(0.7, 5.0, 3)))
    for _ in range(20):  # slivers: one edge ~1e-5 long
        a = rng.uniform(-0.8, 0.8, 2)
        b = a + rng.normal(size=2) * 1e-5
        cc = a + rng.normal(size=2) * 0.3
        add(a, b, cc)
        add(a, cc, b)
    for _ in range(10):  # needles
        a = rng.uniform(-0.8, 0.8, 2)
        d = rng.normal(size=2)
        add(a, a + d * 0.5, a + d * 0.5 + np.array([-d[1], d[0]]) * 1e-4)
        add(a, a + d * 0.5 + np.array([-d[1], d[0]]) * 1e-4, a + d * 0.5)
    for i in (5, 20, 33):  # vertices exactly on pixel centres, duplicated faces (z ties)
        p0, p1, p2 = (c(i), c(i)), (c(i + 6), c(i)), (c(i), c(i + 6))
        add(p0, p1, p2)
        add(p0, p1, p2)
        add(p0, p2, p1)
    add((-6.0, -6.0), (6.5, -6.0), (0.0, 7.0), z=(3.0, 3.0, 3.0))   # huge, |ndc| > 4 (not fast)
    add((-6.0, -6.0), (0.0, 7.0), (6.5, -6.0), z=(3.0, 3.0, 3.0))
    add((1.5, 1.5), (1.7, 1.5), (1.5, 1.8))                           # off screen
    add((c(10) + 1e-4, c(10) + 1e-4), (c(10) + 3e-4, c(10) + 1e-4), (c(10) + 1e-4, c(10) + 3e-4))  # sub-pixel
    add((c(10) + 1e-4, c(10) + 1e-4), (c(10) + 1e-4, c(10) + 3e-4), (c(10) + 3e-4, c(10) + 1e-4))
    # one / two vertices nearer than z_clip: cut by clip_faces into two neighbouring triangles / one triangle
    add((-0.6, -0.6), (0.0, 0.5), (0.6, -0.6), z=(2.0, 0.2, 2.0))
    add((-0.3, 0.2), (0.1, 0.7), (0.5, 0.1), z=(0.3, 1.5, 0.45))
    add((0.2, -0.7), (0.5, -0.2), (0.8, -0.8), z=(0.45, 3.0, 3.0))
    add((-0.8, 0.1), (-0.6, 0.6), (-0.2, 0.2), z=(1.0, 0.49, 0.9))
    # wholly nearer than z_clip = znear/2: removed by clip_faces (they would cover half the image otherwise)
    add((-0.9, -0.9), (0.9, -0.9), (0.0, 0.9), z=(0.3, 0.45, 0.4))
    add((-0.5, -0.5), (0.5, -0.5), (0.0, 0.5), z=(0.2, 0.2, 0.2))
    tris = np.asarray(tris, np.float32)
    verts = tris.reshape(-1, 3)
    faces = np.arange(len(verts), dtype=np.int32).reshape(-1, 3)
    half = len(faces) // 2
    sc = pack_scene([(verts[:half * 3], faces[:half]), (verts[half * 3:], faces[half:] - half * 3)])
    R = np.eye(3, dtype=np.float32)
    T = np.zeros(3, np.float32)
    C = np.zeros(3, np.float32)
    ref = oracle.render_scene(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, S, C, R, T)
    for exact in (False, True):
        eng = OcclusionEngine(sc, 1, RasterConfig(image_size=S, debug_exact=exact), debug_outputs=True)
        Rt, Tt, Ct = (torch.tensor(x[None], device="cuda").contiguous() for x in (R, T, C))
        eng.render(Rt, Tt, Ct)
        assert not (int(eng.status[0]) & (1 | 4))
        assert ref.zclip_straddle
        assert np.array_equal(eng.nhits[0].cpu().numpy(), ref.nhits), exact
        assert np.array_equal(eng.pix_to_face[0].cpu().numpy(), ref.pix_to_face), exact
        assert np.array_equal(eng.obs[0, 3].cpu().numpy(), ref.obs[3]), exact
        assert np.array_equal(eng.n_covered[0].cpu().numpy(), ref.n_covered)
        assert np.array_equal(eng.n_visible[0].cpu().numpy(), ref.n_visible)
        np.testing.assert_allclose(eng.alphas[0].cpu().numpy(), ref.alphas, rtol=RTOL, atol=ATOL_A)
        np.testing.assert_allclose(eng.obs[0, :3].cpu().numpy(), ref.obs[:3], rtol=RTOL, atol=1e-6)


def test_dataset_generator_writes_reference_layout(cuda_lib, tmp_path):
    """datasetGenerator.py:68-124 batched: run folders, 3 images per frame, params rows with finite gradients."""
    import pickle

    import cv2

    from occlusionenv_b200.datasetGenerator import generate_dataset
    n = generate_dataset(str(tmp_path / "Dataset"), num_obj=3, num_frame=4, img_size=64, batch=2, azimuth=1.5, seed=1)
    assert n == 12
    for i in range(3):
        run = tmp_path / "Dataset" / f"run_{i}"
        params = pickle.load(open(run / "params.pickle", "rb")).reshape(-1, 5)
        assert params.shape == (4, 5) and np.array_equal(params[:, 0], np.arange(4)) and np.isfinite(params).all()
        assert (np.abs(params[:, 2] - 1.5) < 0.3).all()
        for j in range(4):
            assert cv2.imread(str(run / "RGB" / f"{j}.jpg")).shape == (64, 64, 3)
            occl = cv2.imread(str(run / "Occl" / f"{j}.png"), cv2.IMREAD_UNCHANGED)
            depth = cv2.imread(str(run / "Depth" / f"{j}.png"), cv2.IMREAD_UNCHANGED)
            assert occl.shape == (64, 64) and depth.shape == (64, 64) and depth.max() > 40


def test_compile_time_tiles_match_oracle_at_128(oracle, cuda_lib):
    """The three compile-time tiles (32x32 default, 32x16 for 3-4 objects, 128x4 for dense meshes) on the same frame."""
    from occlusionenv_b200.engine import OcclusionEngine
    sc = default_scene("teapot")
    S = 128
    _, _, C, R, T = oracle.pose_step(np.zeros(2, np.float32), 0.15, 1.5, 4.0)
    ref = oracle.render_scene(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, S, C, R, T)
    Rt, Tt, Ct = (torch.tensor(x[None], device="cuda").contiguous() for x in (R, T, C))
    for tw, th in [(32, 32), (32, 16), (128, 4)]:
        for dbg in (True, False):
            eng = OcclusionEngine(sc, 1, RasterConfig(image_size=S, tile_w=tw, tile_h=th), debug_outputs=dbg)
            eng.render(Rt, Tt, Ct)
            eng.check_status()
            assert np.array_equal(eng.obs[0, 3].cpu().numpy(), ref.zbuf), (tw, th, dbg)
            np.testing.assert_allclose(eng.obs[0, 0].cpu().numpy(), ref.obs[0], rtol=RTOL, atol=1e-6)
            np.testing.assert_allclose(eng.occl[0].cpu().numpy(), ref.occl, rtol=RTOL, atol=4e-6)
            np.testing.assert_allclose(float(eng.loss[0]), float(ref.loss), rtol=RTOL, atol=1e-6)
            assert np.array_equal(eng.n_covered[0].cpu().numpy(), ref.n_covered)
            assert np.array_equal(eng.n_visible[0].cpu().numpy(), ref.n_visible)
            if dbg:
                assert np.array_equal(eng.pix_to_face[0].cpu().numpy(), ref.pix_to_face), (tw, th)
                assert np.array_equal(eng.nhits[0].cpu().numpy(), ref.nhits), (tw, th)
                np.testing.assert_allclose(eng.alphas[0].cpu().numpy(), ref.alphas, rtol=RTOL, atol=ATOL_A)
