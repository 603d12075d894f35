"""GPU parity: libocclb200.so (through the C-ABI) vs the CPU oracle on identical cameras.

Bit-exact: pix_to_face, pixel counts, soft-hit counts.  fp32 tolerance (stated per assert): silhouettes,
depth, shaded colour, occlusion map, loss."""
import numpy as np
import pytest
import torch

from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.meshes import default_scene

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north-star tolerance for silhouettes / depth / reward
# alpha = 1 - prod_{k<=100}(1 - p_k) is formed in fp32: every factor and every multiply carries an error of
# the order of ulp(1.0) = 1.2e-7 whatever the size of alpha (and the order of the factors is not the
# reference's either: torch.prod vs. shared-memory accumulation), so up to ~16 ulp(1.0) absolute after 100
# factors; the 1e-5 relative bound is what holds wherever alpha >= 0.2.
ATOL_A = 2e-6


def _poses(oracle, kind):
    out = []
    if kind == "step":
        for az, el in [(1.5, 0.0), (np.pi / 2, 0.0), (1.2, 0.25), (1.9, -0.3), (0.7, 0.3), (0.0, 0.0)]:
            _, _, C, R, T = oracle.pose_step(np.zeros(2, np.float32), el, az, 4.0)
            out.append((C, R, T))
    else:
        for az, el in [(1.5, 0.0), (1.0, 0.4), (2.2, -0.2)]:
            out.append(oracle.pose_lookat(4.0, el, az))
    return out


def _render_oracle(oracle, sc, S, pose):
    C, R, T = pose
    return oracle.render_scene(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, S, C, R, T)


@pytest.mark.parametrize("occluder", ["teapot", "box"])
@pytest.mark.parametrize("S", [128, 64])
def test_render_matches_oracle(oracle, cuda_lib, occluder, S):
    from occlusionenv_b200.engine import OcclusionEngine
    sc = default_scene(occluder)
    poses = _poses(oracle, "step") + _poses(oracle, "lookat")
    n = len(poses)
    eng = OcclusionEngine(sc, n, RasterConfig(image_size=S), debug_outputs=True)
    R = torch.tensor(np.stack([p[1] for p in poses]), dtype=torch.float32, device="cuda").contiguous()
    T = torch.tensor(np.stack([p[2] for p in poses]), dtype=torch.float32, device="cuda").contiguous()
    C = torch.tensor(np.stack([p[0] for p in poses]), dtype=torch.float32, device="cuda").contiguous()
    eng.render(R, T, C)
    torch.cuda.synchronize()
    status = eng.status.cpu().numpy()
    assert not (status & (1 | 4 | 8)).any(), status
    for e, pose in enumerate(poses):
        ref = _render_oracle(oracle, sc, S, pose)
        p2f = eng.pix_to_face[e].cpu().numpy()
        assert np.array_equal(p2f, ref.pix_to_face), f"pix_to_face differs at {np.argwhere(p2f != ref.pix_to_face)[:5]}"
        assert np.array_equal(eng.nhits[e].cpu().numpy(), ref.nhits)
        assert np.array_equal(eng.n_covered[e].cpu().numpy(), ref.n_covered)
        assert np.array_equal(eng.n_visible[e].cpu().numpy(), ref.n_visible)
        obs = eng.obs[e].cpu().numpy()
        assert np.array_equal(obs[3], ref.obs[3]), "depth channel must be bit-identical (same fp32 ops)"
        np.testing.assert_allclose(obs[:3], ref.obs[:3], rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(eng.bary[e].cpu().numpy(), ref.bary, rtol=0, atol=0)
        np.testing.assert_allclose(eng.alphas[e].cpu().numpy(), ref.alphas, rtol=RTOL, atol=ATOL_A)
        np.testing.assert_allclose(eng.occl[e].cpu().numpy(), ref.occl, rtol=2 * RTOL, atol=2 * ATOL_A)
        np.testing.assert_allclose(float(eng.loss[e]), float(ref.loss), rtol=RTOL, atol=1e-6)
        if (ref.nhits > 100).any():
            assert status[e] & 2
