"""Diagnostic run on the GPU box: render a few poses, print mismatch statistics against the oracle."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from occlusionenv_b200.config import RasterConfig  # noqa: E402
from occlusionenv_b200.engine import OcclusionEngine  # noqa: E402
from occlusionenv_b200.meshes import default_scene  # noqa: E402
from oracle import oracle as O  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for occ in ("teapot", "box"):
    sc = default_scene(occ)
    poses = []
    for az, el in [(1.5, 0.0), (1.2, 0.25), (0.0, 0.0)]:
        _, _, C, R, T = O.pose_step(np.zeros(2, np.float32), el, az, 4.0)
        poses.append((C, R, T))
    n = len(poses)
    eng = OcclusionEngine(sc, n, RasterConfig(image_size=S), debug_outputs=True)
    R = torch.tensor(np.stack([p[1] for p in poses]), device="cuda").contiguous()
    T = torch.tensor(np.stack([p[2] for p in poses]), device="cuda").contiguous()
    C = torch.tensor(np.stack([p[0] for p in poses]), device="cuda").contiguous()
    eng.render(R, T, C)
    torch.cuda.synchronize()
    print(occ, "status", eng.status.cpu().numpy(), "loss", eng.loss.cpu().numpy())
    for e, (Cc, Rr, Tt) in enumerate(poses):
        ref = O.render_scene(sc.verts, sc.faces, sc.obj_face_start, sc.obj_vert_start, S, Cc, Rr, Tt)
        p2f = eng.pix_to_face[e].cpu().numpy()
        print(f" env{e}: p2f mismatches {(p2f != ref.pix_to_face).sum()} nhits mism {(eng.nhits[e].cpu().numpy() != ref.nhits).sum()}"
              f" ncov {eng.n_covered[e].cpu().numpy()} vs {ref.n_covered} nvis {eng.n_visible[e].cpu().numpy()} vs {ref.n_visible}")
        a = eng.alphas[e].cpu().numpy()
        print(f"   alpha maxabs {np.abs(a - ref.alphas).max():.3e} occl maxabs {np.abs(eng.occl[e].cpu().numpy() - ref.occl).max():.3e}"
              f" loss {float(eng.loss[e]):.6f} vs {float(ref.loss):.6f} depth mism {(eng.obs[e,3].cpu().numpy() != ref.obs[3]).sum()}"
              f" rgb maxabs {np.abs(eng.obs[e,:3].cpu().numpy() - ref.obs[:3]).max():.3e} px>100 {(ref.nhits > 100).sum()}")

# quick throughput probe
sc = default_scene("teapot")
N = 1024
eng = OcclusionEngine(sc, N, RasterConfig(image_size=128))
g = torch.Generator().manual_seed(0)
az = (np.pi / 2 - 0.6) + 1.2 * torch.rand(N, generator=g)
el = -0.3 + 0.6 * torch.rand(N, generator=g)
eng.reset(radius=4.0, azimuth=az, elevation=el)
act = torch.randn(N, 2, generator=torch.Generator().manual_seed(1)).cuda()
for _ in range(3):
    eng.step(act)
torch.cuda.synchronize()
t = time.time()
for _ in range(10):
    eng.step(act)
torch.cuda.synchronize()
dt = (time.time() - t) / 10
print(f"probe: N={N} step {dt*1e3:.3f} ms -> {N/dt:.0f} env-steps/s; status or {int(eng.status.max())}")
