"""Small end-to-end cases for compute-sanitizer (one tool per run):
   compute-sanitizer --tool memcheck  python tools/sanitize_case.py
   compute-sanitizer --tool racecheck python tools/sanitize_case.py
(compute-sanitizer is closed on some GPU pools: the script also runs plain, as a smoke of every instantiation.)
Covers the debug and the production instantiations (32x32, 32x16 with the evaluate-once path, 128x4 on a dense mesh),
the differentiable kernels, the masked reset list kernel, cut faces (clip kernel), the chunked workspace (two lanes)
and the two-plane observation layout."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from occlusionenv_b200.config import RasterConfig
from occlusionenv_b200.engine import OcclusionEngine
from occlusionenv_b200.meshes import default_scene, procedural_scene

heavy = len(sys.argv) < 2 or sys.argv[1] != "light"


def drive(tag, eng, az, el, act, radius=4.0):
    eng.reset(radius=radius, azimuth=az, elevation=el)
    eng.step(act, with_grad=True)
    eng.step(act, with_grad=False)
    m = torch.zeros(eng.n, dtype=torch.uint8, device="cuda")
    m[::2] = 1
    eng.reset(radius=radius, azimuth=0.0, elevation=0.0, mask=m)
    torch.cuda.synchronize()
    print(tag, "status_or", eng.check_status(raise_on=0), "loss", eng.loss.cpu().numpy()[:3], flush=True)


az3, el3 = torch.tensor([1.5, 1.2, 0.0]), torch.tensor([0.0, 0.2, 0.0])
act3 = torch.tensor([[0.3, -1.0], [1.0, 1.0], [0.0, 0.0]], device="cuda")
for occ, S in (("teapot", 64), ("box", 48)):
    drive(f"debug {occ} {S}", OcclusionEngine(default_scene(occ), 3, RasterConfig(image_size=S), debug_outputs=True), az3, el3, act3)
# production instantiations
drive("production 32x32", OcclusionEngine(default_scene("box"), 3, RasterConfig(image_size=128)), az3, el3, act3)
drive("two planes", OcclusionEngine(default_scene("box"), 3, RasterConfig(image_size=64, obs_planes=2)), az3, el3, act3)
drive("chunked, two lanes", OcclusionEngine(default_scene("teapot"), 3, RasterConfig(image_size=64, ws_budget_mb=1)), az3, el3, act3)
# faces cut at z_clip (camera close to the teapot): clip kernel, forward and differentiable
drive("cut faces", OcclusionEngine(default_scene("teapot"), 3, RasterConfig(image_size=64)), torch.tensor([1.5, 1.4, 1.6]),
      torch.tensor([0.3, 0.2, 0.1]), act3, radius=1.8)
if heavy:
    # three objects: the 32x16 tile with the evaluate-once path; dense meshes: the 128x4 tile, K = 100 live
    sc3 = procedural_scene(2, n_obj=3, subdiv=3)
    drive("three objects 32x16", OcclusionEngine(sc3, 3, RasterConfig(image_size=64)), torch.tensor([-0.35, 0.3, 0.0]),
          torch.tensor([0.1, 0.1, 0.1]), act3)
    dense = [procedural_scene(s, n_obj=3, subdiv=5) for s in (2, 3)]
    e = OcclusionEngine(None, 2, RasterConfig(image_size=128), per_env_scenes=dense)
    print("dense tile", int(e.c.tile_w), int(e.c.tile_h))
    drive("dense 128x4", e, torch.tensor([-0.35, 0.3]), torch.tensor([0.1, 0.1]), act3[:2].contiguous())
print("sanitize case done")
