"""Config-3 workload statistics (one env): hits per pixel, K-overflow slots per tile."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.engine import OcclusionEngine
from occlusionenv_b200.meshes import procedural_scene

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for seed, az in ((0, -0.3), (1, 0.2), (2, 0.45)):
    sc = procedural_scene(seed, n_obj=3, subdiv=5)
    eng = OcclusionEngine(sc, 1, RasterConfig(image_size=S), debug_outputs=True)
    eng.reset(radius=4.0, azimuth=az, elevation=0.1)
    torch.cuda.synchronize()
    nh = eng.nhits[0].cpu().numpy()          # (n_obj, S, S)
    ov = nh > 100
    T = S // 32
    per_tile = ov.reshape(3, T, 32, T, 32).sum((0, 2, 4))
    hits_tile = nh.reshape(3, T, 32, T, 32).sum((0, 2, 4))
    cov = (nh > 0).reshape(3, -1).sum(1)
    print(f"seed {seed}: touched px/obj {cov.tolist()} total hits {nh.sum()} overflow slots {ov.sum()} hits in overflow {nh[ov].sum()} "
          f"mean hits/ovf slot {nh[ov].mean():.0f} max {nh.max()} | tiles with overflow {(per_tile>0).sum()} max ovf slots/tile {per_tile.max()} "
          f"max hits/tile {hits_tile.max()} status {int(eng.status[0])}")
