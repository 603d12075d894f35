"""Oracle parity AT THE BENCHMARKED SHAPES, on the PRODUCTION kernel instantiations (no debug outputs):

  * C2 / C4: bench.py's own workload (teapot + box occluder, 4096 envs, 128x128, bench.make_poses), one reset + one
    step of the whole batch through occl_step -- ``raster_kernel<0|1,32,32,0>``, the kernel behind the headline
    number -- then 32 seeded random envs are compared with the CPU oracle's state machine
    (``environment.py:352-396`` restated in oracle/oracle.py);
  * C3: per-env procedural meshes of three 20 480-face objects at 256x256, ``raster_kernel<0,128,4,0>`` (the dense
    production tile), 2 envs against the oracle;
  * the chunked workspace (OcclConfig.ws_budget_mb) gives the same results as the unchunked one.
Counts / done / depth bit-exact; observation, occlusion map, loss, reward at the tolerances of test_gpu_step.py.
"""
import math
import os
import sys

import numpy as np
import pytest
import torch

from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.meshes import default_scene, procedural_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
RTOL = 1e-5
ATOL_A = 2e-6


def _compare_env(oracle, sc, S, eng, e, az0, el0, act, check_grad_with=None):
    ref = oracle.OracleOcclusionEnv(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, img_size=S)
    ref.reset(radius=4.0, azimuth=az0, elevation=el0)
    obs, rew, done, info = ref.step(act)
    out = ref.last
    assert np.array_equal(eng.n_covered[e].cpu().numpy(), out.n_covered), e
    assert np.array_equal(eng.n_visible[e].cpu().numpy(), out.n_visible), e
    assert bool(eng.done[e]) == bool(done), e
    got = eng.obs[e].cpu().numpy()
    assert np.array_equal(got[3], obs[0][3]), f"depth of env {e} is not bit-exact"
    np.testing.assert_allclose(got[:3], obs[0][:3], rtol=RTOL, atol=1e-6, err_msg=f"env {e}")
    np.testing.assert_allclose(eng.occl[e].cpu().numpy(), info["full_state"], rtol=2 * RTOL, atol=2 * ATOL_A, err_msg=f"env {e}")
    np.testing.assert_allclose(float(eng.loss[e]), float(info["full_reward"]), rtol=RTOL, atol=1e-6, err_msg=f"env {e}")
    np.testing.assert_allclose(float(eng.reward[e]), float(rew), rtol=RTOL, atol=2e-6, err_msg=f"env {e}")
    np.testing.assert_array_equal(eng.position[e].cpu().numpy(), info["position"])
    assert float(eng.elevation[e]) == float(ref.elevation) and float(eng.azimuth[e]) == float(ref.azimuth)
    return ref


@pytest.mark.parametrize("grad", [False, True], ids=["c2_fwd", "c4_grad"])
def test_bench_workload_sampled_envs_match_oracle(oracle, cuda_lib, grad):
    import bench
    from occlusionenv_b200.engine import OcclusionEngine
    from oracle import dense_torch as D
    N, S = 4096, 128
    sc = default_scene("box")
    eng = OcclusionEngine(sc, N, RasterConfig(image_size=S))       # debug_outputs=False: the production instantiation
    assert (int(eng.c.tile_w), int(eng.c.tile_h)) == (32, 32)
    az, el, actions = bench.make_poses(N, 0)
    eng.reset(radius=4.0, azimuth=az, elevation=el)
    prev = eng.full_reward.cpu().numpy().copy()
    mass = eng.object_mass.cpu().numpy().copy()
    eng.step(actions[0].cuda(), with_grad=grad)
    torch.cuda.synchronize()
    assert not (eng.check_status(raise_on=0) & (1 | 4 | 8))
    ids = np.random.default_rng(2026).choice(N, size=32, replace=False)
    for k, e in enumerate(sorted(int(i) for i in ids)):
        _compare_env(oracle, sc, S, eng, e, float(az[e]), float(el[e]), actions[0][e].numpy())
        if grad and k < 3:  # float64 autograd of the dense formulation is minutes per env at 128^2: three envs
            _, _, g, _ = D.reward_and_grad(sc, S, actions[0][e].numpy().astype(np.float64), float(el[e]), float(az[e]), 4.0,
                                           float(prev[e]), float(mass[e]), float(oracle.PROJ_SCALE),
                                           float(oracle.BLUR_RADIUS), float(oracle.SIGMA))
            scale = max(np.abs(g).max(), 1e-6)
            assert np.abs(eng.grad_action[e].cpu().numpy() - g).max() <= 1e-3 * scale, (e, eng.grad_action[e], g)
    if grad:
        assert torch.isfinite(eng.grad_action).all()


def test_config3_dense_meshes_production_tile_match_oracle(oracle, cuda_lib):
    from occlusionenv_b200.engine import OcclusionEngine
    S, n = 256, 2
    scenes = [procedural_scene(s, n_obj=3, subdiv=5) for s in (2, 3)]  # seeds whose layout shows occlusion at these poses
    assert scenes[0].faces.shape[0] == 3 * 20480
    eng = OcclusionEngine(None, n, RasterConfig(image_size=S), per_env_scenes=scenes)  # production instantiation
    assert (int(eng.c.tile_w), int(eng.c.tile_h)) == (128, 4)
    az0 = np.array([-0.35, 0.3], np.float32)
    eng.reset(radius=4.0, azimuth=torch.tensor(az0), elevation=0.1)
    act = np.array([[1.0, 0.5], [-0.3, 0.9]], np.float32)
    eng.step(torch.tensor(act, device="cuda"))
    torch.cuda.synchronize()
    st = eng.check_status(raise_on=0)
    assert st & 2, "K = 100 must be live on these meshes"
    assert not (st & (1 | 4 | 8)), st
    for e, sc in enumerate(scenes):
        _compare_env(oracle, sc, S, eng, e, float(az0[e]), 0.1, act[e])


def test_chunked_workspace_equals_unchunked(cuda_lib):
    """OcclConfig.ws_budget_mb: 96 envs rasterised in chunks of a few envs give exactly the unchunked results."""
    from occlusionenv_b200.engine import OcclusionEngine
    sc = default_scene("teapot")
    N, S = 96, 64
    g = torch.Generator().manual_seed(3)
    az = (math.pi / 2 - 0.6) + 1.2 * torch.rand(N, generator=g)
    el = -0.3 + 0.6 * torch.rand(N, generator=g)
    act = torch.randn(N, 2, generator=g).cuda()
    res = []
    for budget in (0, 4):
        eng = OcclusionEngine(sc, N, RasterConfig(image_size=S, ws_budget_mb=budget), debug_outputs=True)
        eng.reset(radius=4.0, azimuth=az, elevation=el)
        eng.step(act, with_grad=True)
        eng.check_status()
        res.append(eng)
    a, b = res
    assert b.workspace.numel() < a.workspace.numel() // 4
    for name in ("obs", "occl", "alphas", "pix_to_face", "nhits", "n_covered", "n_visible", "reward", "done", "loss", "status"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    # the tangent sums are float adds in arrival order: equal up to that noise
    torch.testing.assert_close(a.grad_action, b.grad_action, rtol=1e-4, atol=1e-6)


def test_auto_reset_keeps_terminal_info(oracle, cuda_lib):
    """infos of an env that finished AND was auto-reset in the same call are the terminal step's
    (SubProcVecEnv.py:210-214): full_state, full_reward, counts, position; obs is the reset observation."""
    from occlusionenv_b200.SubProcVecEnv import BatchedOcclusionVecEnv
    sc = default_scene("teapot")
    S, n = 32, 4
    venv = BatchedOcclusionVecEnv(n, data=None, img_size=S, keep_terminal_obs=True)
    az0 = np.array([0.0, 1.5, 0.02, 1.4], np.float32)  # envs 0 and 2 are unoccluded: done on the first step
    venv.reset(azimuth=torch.tensor(az0))
    a = np.array([[0.1, 0.2], [0.3, -0.2], [-0.5, 0.1], [0.2, 0.2]], np.float32)
    obs, rews, dones, infos = venv.step(torch.tensor(a))
    assert dones.cpu().tolist() == [True, False, True, False]
    for e in range(n):
        ref = oracle.OracleOcclusionEnv(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, img_size=S)
        ref.reset(azimuth=az0[e])
        o, r, d, info = ref.step(a[e])
        got = infos[e]
        np.testing.assert_allclose(got["full_state"][0, ..., 3].cpu().numpy(), info["full_state"], rtol=2 * RTOL, atol=2 * ATOL_A)
        np.testing.assert_allclose(float(got["full_reward"]), float(info["full_reward"]), rtol=RTOL, atol=1e-6)
        np.testing.assert_array_equal(got["position"].cpu().numpy(), info["position"])
        assert np.array_equal(got["n_covered"].cpu().numpy(), ref.last.n_covered)
        if d:
            np.testing.assert_allclose(got["terminal_observation"][0].cpu().numpy(), o[0], rtol=RTOL, atol=1e-6)
            o = ref.reset()
        np.testing.assert_allclose(obs[e].cpu().numpy(), o[0], rtol=RTOL, atol=1e-6)
    assert venv.check_status() & ~(2 | 16) == 0


def test_scene_sampler_reset_redraws_unoccluded_envs(cuda_lib):
    """new_scene semantics of OcclusionEnv.reset (environment.py:292-298,327-328) in the batched env: every env gets
    its own sampled scene; envs whose reset shows no occlusion are redrawn (at most 10 renders)."""
    from occlusionenv_b200.SubProcVecEnv import BatchedOcclusionVecEnv
    calls = []

    def sampler():
        # odd draws put the second object far to the side (no occlusion at azimuth 0), even draws right behind the first
        k = len(calls)
        calls.append(k)
        sc = procedural_scene(k, n_obj=2, subdiv=2)
        v = sc.verts.copy()
        v0 = sc.obj_vert_start[1]
        v[v0:] += np.float32([(0.0 if k % 2 == 0 else 6.0), 0.0, 0.0]) - v[v0:].mean(0) + np.float32([0.0, 0.0, 1.0])
        sc.verts = np.ascontiguousarray(v, np.float32)
        return sc

    n, S = 6, 32
    venv = BatchedOcclusionVecEnv(n, img_size=S, scene_sampler=sampler)
    n0 = len(calls)
    assert n0 == n
    venv.reset(azimuth=0.0)
    assert len(calls) > n0 + n, "unoccluded envs must have been redrawn"
    assert bool((venv.engine.full_reward > 0.1).all())
    obs, rews, dones, infos = venv.step(torch.zeros(n, 2))
    assert obs.shape == (n, 4, S, S)


def test_full_batch_properties_determinism_and_env_permutation(cuda_lib):
    """Size-independent properties at the full config-2 shape (4096 envs x 128^2, production kernel):
    (a) the forward transition is run-to-run deterministic bit for bit (integer accumulators: no float atomics);
    (b) envs do not see each other: permuting the poses / actions over the batch permutes every output bit for bit
        (the same env lands in another CTA, another SM, another position of the launch);
    (c) a checksum of checksums: the batch loss sum equals the sum over the permuted batch exactly (float64 of fp32)."""
    import bench
    from occlusionenv_b200.engine import OcclusionEngine
    N, S = 4096, 128
    eng = OcclusionEngine(default_scene("box"), N, RasterConfig(image_size=S))
    az, el, actions = bench.make_poses(N, 0)
    a = actions[0].cuda()

    def run(az_, el_, a_):
        eng.reset(radius=4.0, azimuth=az_, elevation=el_)
        eng.step(a_)
        torch.cuda.synchronize()
        return [t.clone() for t in (eng.obs, eng.occl, eng.reward, eng.loss, eng.done, eng.n_covered, eng.n_visible,
                                    eng.position)]

    first = run(az, el, a)
    again = run(az, el, a)
    names = ("obs", "occl", "reward", "loss", "done", "n_covered", "n_visible", "position")
    for name, x, y in zip(names, first, again):
        nd = int((x != y).sum())
        assert nd == 0, f"the forward transition is not deterministic: {name} differs in {nd} values, max |d| {float((x.float() - y.float()).abs().max()):.3e}"
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(7))
    shuffled = run(az[perm], el[perm], a[perm.cuda()])
    for name, x, y in zip(names, first, shuffled):
        nd = int((x[perm.to(x.device)] != y).sum())
        assert nd == 0, f"an env's result depends on its position in the batch: {name} differs in {nd} values"
    assert float(first[3].double().sum()) == pytest.approx(float(shuffled[3].double().sum()), rel=0, abs=0)
    assert float(first[3].max()) > 1.0  # (not the trivial all-background batch)


def test_full_batch_kernel_instantiations_agree_bit_for_bit(cuda_lib):
    """The same transition through different instantiations of the tile rasteriser, at the full config-2 shape:
    (a) the differentiable kernel (rounds of 128 faces, 3 CTAs/SM) writes the forward outputs of the forward kernel;
    (b) a masked reset (device-side env list + persistent ``raster_list_kernel``) gives the flagged envs what a reset of
        the whole batch gives them, and leaves every other env alone;
    (c) the compact transport layout (``obs_planes=2``: grey + depth) is planes 0 and 3 of the reference layout;
    (d) a generic-tile engine (32x32 passed at run time to the non-specialised kernel) agrees with the compile-time one."""
    import bench
    from occlusionenv_b200.engine import OcclusionEngine
    N, S = 4096, 128
    sc = default_scene("box")
    az, el, actions = bench.make_poses(N, 0)
    a = actions[0].cuda()
    names = ("obs", "occl", "reward", "loss", "done", "n_covered", "n_visible")

    def outs(e):
        torch.cuda.synchronize()
        return [t.clone() for t in (e.obs, e.occl, e.reward, e.loss, e.done, e.n_covered, e.n_visible)]

    def same(tag, xs, ys, rows=None):
        for name, x, y in zip(names, xs, ys):
            if rows is not None:
                x, y = x[rows], y[rows]
            nd = int((x != y).sum())
            assert nd == 0, f"{tag}: {name} differs in {nd} values"

    eng = OcclusionEngine(sc, N, RasterConfig(image_size=S))
    eng.reset(radius=4.0, azimuth=az, elevation=el)
    base_reset = outs(eng)
    eng.step(a)
    fwd = outs(eng)
    # (a)
    eng.reset(radius=4.0, azimuth=az, elevation=el)
    eng.step(a, with_grad=True)
    same("differentiable vs forward kernel", fwd, outs(eng))
    assert torch.isfinite(eng.grad_action).all()
    # (b) the state is now the one after the step; reset a random third of the envs back to their start pose
    mask = (torch.rand(N, generator=torch.Generator().manual_seed(3)) < 0.33).cuda()
    after_step = outs(eng)
    eng.reset(radius=4.0, azimuth=az, elevation=el, mask=mask)
    got = outs(eng)
    # a reset has no reward, and a masked one leaves `done` of the step in place (the mask may alias it)
    keep = [i for i, n in enumerate(names) if n not in ("done", "reward")]
    same("masked reset, flagged envs", [base_reset[i] for i in keep], [got[i] for i in keep], rows=mask)
    same("masked reset, other envs", [after_step[i] for i in keep], [got[i] for i in keep], rows=~mask)
    del eng
    # (c)
    e2 = OcclusionEngine(sc, N, RasterConfig(image_size=S, obs_planes=2))
    e2.reset(radius=4.0, azimuth=az, elevation=el)
    e2.step(a)
    torch.cuda.synchronize()
    assert e2.obs.shape == (N, 2, S, S)
    assert torch.equal(e2.obs[:, 0], fwd[0][:, 0]) and torch.equal(e2.obs[:, 1], fwd[0][:, 3])
    assert torch.equal(fwd[0][:, 0], fwd[0][:, 1]) and torch.equal(fwd[0][:, 0], fwd[0][:, 2])  # R = G = B
    same("two-plane layout", fwd[1:], [e2.occl, e2.reward, e2.loss, e2.done, e2.n_covered, e2.n_visible])
    del e2
    # (d) 16 x 64 is not a compile-time tile: the generic instantiation
    e3 = OcclusionEngine(sc, N, RasterConfig(image_size=S, tile_w=16, tile_h=64))
    e3.reset(radius=4.0, azimuth=az, elevation=el)
    e3.step(a)
    same("generic 16x64 tile vs compile-time 32x32", fwd, outs(e3))


def test_config3_dense_batch_properties(cuda_lib):
    """Config-3 meshes (3 x 20 480 faces per env, 256^2, 128x4 production tile with the evaluate-once K-overflow path) at a
    batch that fills the GPU several times (192 envs = 24 576 CTAs): the forward transition is run-to-run identical,
    a workspace chunked into a few envs per launch gives the same bits, and the differentiable kernel writes the same
    forward outputs (its K-overflow slots record tangents next to the terms)."""
    from occlusionenv_b200.engine import OcclusionEngine
    S, N = 256, 192
    base = [procedural_scene(s, n_obj=3, subdiv=5) for s in (2, 3, 5)]
    scenes = [base[i % 3] for i in range(N)]
    g = torch.Generator().manual_seed(11)
    az = -0.5 + torch.rand(N, generator=g)
    act = torch.randn(N, 2, generator=g).cuda()
    names = ("obs", "occl", "reward", "loss", "done", "n_covered", "n_visible")

    def run(budget, grad):
        eng = OcclusionEngine(None, N, RasterConfig(image_size=S, ws_budget_mb=budget), per_env_scenes=scenes)
        assert (int(eng.c.tile_w), int(eng.c.tile_h)) == (128, 4)
        out = []
        for _ in range(2):
            eng.reset(radius=4.0, azimuth=az, elevation=0.1)
            eng.step(act, with_grad=grad)
            torch.cuda.synchronize()
            out.append([getattr(eng, n).clone() for n in names])
        st = eng.check_status(raise_on=0)
        assert st & 2 and not (st & (1 | 4 | 8)), st
        return out, eng.workspace.numel()

    (first, again), ws_full = run(0, False)
    assert float(first[3].max()) > 1.0
    for n, x, y in zip(names, first, again):
        assert torch.equal(x, y), f"config-3 forward transition is not deterministic: {n}"
    (chunked, _), ws_small = run(128, False)
    assert ws_small < ws_full
    for n, x, y in zip(names, first, chunked):
        assert torch.equal(x, y), f"chunked workspace differs: {n}"
    (gradf, _), _ = run(0, True)
    for n, x, y in zip(names, first, gradf):
        assert torch.equal(x, y), f"differentiable kernel's forward outputs differ: {n}"


def test_incremental_observation_delivery_is_lossless(cuda_lib):
    """OcclOutputs.obs_tile_state: tiles that were background at the last render into a destination and are background
    now are not stored again.  After a reset, steps that move the objects across tiles, a masked reset and a
    differentiable step the destination is bit-identical to the fully written one -- also for the two-plane layout, an
    external destination tensor and a chunked workspace -- and tiles really are skipped (a poisoned skip survives)."""
    from occlusionenv_b200.engine import OcclusionEngine
    N, S = 96, 128
    sc = default_scene("box")
    g = torch.Generator().manual_seed(5)
    az = (math.pi / 2 - 0.9) + 1.8 * torch.rand(N, generator=g)
    el = -0.5 + torch.rand(N, generator=g)
    acts = [torch.randn(N, 2, generator=g).cuda() for _ in range(6)]
    mask = (torch.rand(N, generator=g) < 0.4).cuda()
    for planes, budget in ((4, 0), (2, 0), (4, 8)):
        cfg = RasterConfig(image_size=S, obs_planes=planes, ws_budget_mb=budget)
        full = OcclusionEngine(sc, N, cfg)
        inc = OcclusionEngine(sc, N, cfg)
        dest = torch.full((N, planes, S, S), 7.0, device="cuda")     # an external destination with foreign content
        state = inc.incremental_obs(dest)
        assert state.shape == (N, 16)

        def both(fn):
            fn(full, None)
            fn(inc, dest)
            torch.cuda.synchronize()
            assert torch.equal(full.obs, dest), (planes, budget)
            for name in ("occl", "loss", "done", "n_covered", "n_visible"):
                assert torch.equal(getattr(full, name), getattr(inc, name)), name

        both(lambda e, d: e.reset(radius=4.0, azimuth=az, elevation=el, obs=d))
        for k, a in enumerate(acts):
            both(lambda e, d: e.step(a, with_grad=(k == 3), obs=d))
            if k == 2:
                both(lambda e, d: e.reset(radius=4.0, azimuth=0.0, elevation=0.0, mask=mask, obs=d))
        # tiles are really skipped: poison the destination where the state says "background before and now"
        skip = state[:, 8].clone()                                    # tiles 0..31 (a 128^2 image has 16)
        assert int(skip.ne(0).sum()) > 0
        e0 = int(torch.nonzero(skip)[0])
        t0 = int(torch.nonzero(torch.tensor([(int(skip[e0]) >> b) & 1 for b in range(16)]))[0])
        ty, tx = divmod(t0, S // 32)
        inc.step(acts[0], obs=dest)
        full.step(acts[0])
        torch.cuda.synchronize()
        assert torch.equal(full.obs, dest)
        poisoned = bool((int(state[e0, 8]) >> t0) & 1)
        if poisoned:
            dest[e0, 0, ty * 32, tx * 32] = 123.0
        inc.step(-acts[0], obs=dest)
        full.step(-acts[0])
        torch.cuda.synchronize()
        if poisoned and ((int(state[e0, 8]) >> t0) & 1):
            assert float(dest[e0, 0, ty * 32, tx * 32]) == 123.0   # background before and now: not rewritten
        # ... and starting over restores a full write
        inc.incremental_obs(dest)
        dest.fill_(9.0)
        inc.step(acts[1], obs=dest)
        full.step(acts[1])
        torch.cuda.synchronize()
        assert torch.equal(full.obs, dest)
