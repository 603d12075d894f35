"""Odd-configuration probes: 512^2 batch throughput, n_obj = 1 and 4, default OcclusionEnv size."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.engine import OcclusionEngine
from occlusionenv_b200.meshes import default_scene, load_teapot, make_box, pack_scene

def run(sc, N, S, tag):
    eng = OcclusionEngine(sc, N, RasterConfig(image_size=S))
    g = torch.Generator().manual_seed(0)
    az = (np.pi / 2 - 0.6) + 1.2 * torch.rand(N, generator=g)
    el = -0.3 + 0.6 * torch.rand(N, generator=g)
    eng.reset(radius=4.0, azimuth=az, elevation=el)
    act = torch.randn(N, 2, generator=g).cuda()
    for _ in range(3):
        eng.step(act)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(10):
        eng.step(act)
    torch.cuda.synchronize(); dt = (time.time() - t) / 10
    print(f"{tag}: N={N} S={S} n_obj={sc.n_obj} F={sc.faces.shape[0]} {dt*1e3:.2f} ms/step -> {N/dt:.0f} env-steps/s status_or={int(eng.status.max())} "
          f"loss[:3]={eng.loss[:3].cpu().numpy()}")

tv, tf = load_teapot()
run(default_scene("box"), 512, 512, "box 512^2")
run(default_scene("teapot"), 512, 512, "teapot 512^2")
run(default_scene("teapot"), 1024, 256, "teapot 256^2")
run(pack_scene([(tv, tf)]), 256, 128, "single object")
bv, bf = make_box()
run(pack_scene([(tv, tf), (tv + np.array([2, 0, 0], np.float32), tf), (bv + np.array([0.8, 0, 0], np.float32), bf), (tv + np.array([1, 0, 1.5], np.float32), tf)]), 256, 128, "four objects")
