"""Small end-to-end case for compute-sanitizer (one tool per gpurun call):
   compute-sanitizer --tool memcheck  python tools/sanitize_case.py
   compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.engine import OcclusionEngine
from occlusionenv_b200.meshes import default_scene

for occ, S in (("teapot", 64), ("box", 48)):
    eng = OcclusionEngine(default_scene(occ), 3, RasterConfig(image_size=S), debug_outputs=True)
    eng.reset(radius=4.0, azimuth=torch.tensor([1.5, 1.2, 0.0]), elevation=torch.tensor([0.0, 0.2, 0.0]))
    act = torch.tensor([[0.3, -1.0], [1.0, 1.0], [0.0, 0.0]], device="cuda")
    eng.step(act, with_grad=True)
    eng.step(act, with_grad=False)
    m = torch.tensor([1, 0, 1], dtype=torch.uint8, device="cuda")
    eng.reset(radius=4.0, azimuth=0.0, elevation=0.0, mask=m)
    torch.cuda.synchronize()
    print(occ, S, "status", eng.status.cpu().numpy(), "loss", eng.loss.cpu().numpy())
print("sanitize case done")
