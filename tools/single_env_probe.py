"""Latency of the drop-in single OcclusionEnv (N=1 view of the engine) and of small batches."""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from occlusionenv_b200.environment import OcclusionEnv
from occlusionenv_b200.SubProcVecEnv import BatchedOcclusionVecEnv

for S in (128, 512):
    env = OcclusionEnv(img_size=S)
    env.reset(azimuth=1.5)
    a = torch.tensor([0.3, -1.0])
    for _ in range(20):
        env.step(a)
    torch.cuda.synchronize(); t = time.time()
    n = 200
    for _ in range(n):
        obs, r, d, info = env.step(a)
    torch.cuda.synchronize(); dt = (time.time() - t) / n
    print(f"OcclusionEnv img_size={S}: {dt*1e6:.0f} us/step ({1/dt:.0f} steps/s)")
    ag = torch.nn.Parameter(torch.tensor([0.3, -1.0]))
    torch.cuda.synchronize(); t = time.time()
    for _ in range(50):
        obs, r, d, info = env.step(ag); r.backward()
    torch.cuda.synchronize(); dt = (time.time() - t) / 50
    print(f"  differentiable step + backward: {dt*1e6:.0f} us")
for N in (8, 64, 512):
    v = BatchedOcclusionVecEnv(N, img_size=128)
    v.reset()
    a = torch.randn(N, 2, device="cuda")
    for _ in range(10):
        v.step(a)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(100):
        v.step(a)
    torch.cuda.synchronize(); dt = (time.time() - t) / 100
    print(f"BatchedOcclusionVecEnv N={N} 128^2 (auto-reset on): {dt*1e6:.0f} us/step ({N/dt:.0f} env-steps/s)")
