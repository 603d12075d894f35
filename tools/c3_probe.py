"""Config-3 probe: per-env procedural 20k-face meshes (3 objects), 256x256; times a few steps."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.engine import OcclusionEngine
from occlusionenv_b200.meshes import procedural_scene

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
TW = int(sys.argv[3]) if len(sys.argv) > 3 else 0
TH = int(sys.argv[4]) if len(sys.argv) > 4 else 0
base = [procedural_scene(s, n_obj=3, subdiv=5) for s in range(4)]
scenes = [base[i % 4] for i in range(N)]
eng = OcclusionEngine(None, N, RasterConfig(image_size=S, tile_w=TW, tile_h=TH), per_env_scenes=scenes)
g = torch.Generator().manual_seed(0)
az = -0.5 + torch.rand(N, generator=g)
eng.reset(radius=4.0, azimuth=az, elevation=0.1)
torch.cuda.synchronize()
act = torch.randn(N, 2, generator=g).cuda()
for _ in range(2):
    eng.step(act)
torch.cuda.synchronize()
t = time.time()
K = 5
for _ in range(K):
    eng.step(act)
torch.cuda.synchronize()
dt = (time.time() - t) / K
print(f"C3 probe: tile {eng.c.tile_w}x{eng.c.tile_h} N={N} S={S} faces={eng.c.n_faces} step {dt*1e3:.2f} ms -> {N/dt:.0f} env-steps/s; status {int(eng.status.max())} loss {eng.loss[:4].cpu().numpy()}")
