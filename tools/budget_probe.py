"""Workspace-budget sweep: C2 (4096 envs x 128^2, box) and C3 (8192 envs x 256^2) throughput against OcclConfig.ws_budget_mb
(the chunk of envs per launch; two scratch sets alternate on two streams when there is more than one chunk).
usage: python tools/budget_probe.py [c2 budgets, comma separated] [c3 budgets]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.engine import OcclusionEngine
from occlusionenv_b200.meshes import default_scene

dev = "cuda:0"
h = bench.Harness(1, dev)
c2 = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,256,512,1024").split(",") if x]
c3 = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1024,2048,4096").split(",") if x]
N = 4096
az, el, actions = bench.make_poses(N, 0)
acts = actions.to(dev)
for b in c2:
    eng = OcclusionEngine(default_scene("box"), N, RasterConfig(image_size=128, ws_budget_mb=b), device=dev)
    eng.reset(radius=4.0, azimuth=az, elevation=el)
    step = lambda i: eng.step(acts[i % 8])
    for i in range(10):
        step(i)
    ms = h.timed(step, 60)
    print(f"c2 ws_budget_mb={b} workspace={eng.workspace.numel() / 2**20:.0f} MB -> {N * 60 / ms * 1e3:.0f} env-steps/s", flush=True)
    del eng
    torch.cuda.empty_cache()
if c3:
    scenes = bench.c3_scenes(64)
    n = 8192
    g = torch.Generator().manual_seed(0)
    azr = -0.5 + torch.rand(n, generator=g)
    a = torch.randn(2, n, 2, generator=g)
    acts3 = torch.stack([a[0], -a[0], a[1], -a[1]]).to(dev)
    for b in c3:
        eng = OcclusionEngine(None, n, RasterConfig(image_size=256, ws_budget_mb=b), device=dev, per_env_scenes=scenes,
                              replicate_scenes=True)
        eng.reset(radius=4.0, azimuth=azr, elevation=0.1)
        step = lambda i: eng.step(acts3[i % 4])
        step(0); step(1)
        ms = h.timed(step, 6)
        print(f"c3 ws_budget_mb={b} workspace={eng.workspace.numel() / 2**30:.2f} GB -> {n * 6 / ms * 1e3:.0f} env-steps/s", flush=True)
        del eng
        torch.cuda.empty_cache()
